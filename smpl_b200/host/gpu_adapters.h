// Drop-in adapters: the reference's plugin interfaces implemented over the C ABI.
//
//   GpuCollisionSpace  : sbpl::motion::CollisionChecker        (replaces sbpl::collision::CollisionSpace,
//                         sbpl_collision_checking/src/collision_space.cpp:532-581) + batched entry points
//   GpuRobotModel      : sbpl::motion::ForwardKinematicsInterface (replaces KDLRobotModel's FK + limits,
//                         sbpl_kdl_robot_model/src/kdl_robot_model.cpp:326-423)
//   GpuBfsHeuristic    : sbpl::motion::RobotHeuristic          (replaces BfsHeuristic + BFS_3D,
//                         smpl/src/heuristic/bfs_heuristic.cpp)
// Ownership follows the reference: the caller owns the context and the tables, the adapters hold raw
// non-owning pointers (robot_planning_space.h:68-71, collision_space.h:227).  Errors are `false` / sentinel
// values, never exceptions.  One adapter set per planner thread (the context is not re-entrant).
#ifndef SMPLHOST_GPU_ADAPTERS_H
#define SMPLHOST_GPU_ADAPTERS_H

#include <functional>
#include <string>
#include <vector>

#include "../../include/smplgpu.h"
#include "robot_tables.h"
#ifdef SMPLHOST_REFERENCE_HEADERS
// built inside the reference's tree: the reference's own interface headers
#include <smpl/collision_checker.h>
#include <smpl/robot_model.h>
#include <smpl/heuristic/robot_heuristic.h>
#else
// stand-alone build: the same interfaces restated (names, signatures, defaults as in the reference)
#include "smpl/interfaces.h"
#endif

namespace smplhost {

/// Answers of the device for ONE expansion of the reference's search, shared by the three adapters.
///
/// An unchanged caller (ManipLattice::GetSuccs + ARAStar::expand, manip_lattice.cpp:219-313, arastar.cpp:531-568)
/// asks its plug-ins ~65 questions per expansion, one virtual call each: FK + metric goal distance of the parent,
/// joint limits + edge validity per action, FK per valid successor (isGoal), goal heuristic per new state.  One
/// launch and one PCIe round trip each would make the device SLOWER than the reference's CPU.  The adapters are
/// given the motion-primitive table instead (the same table the action space was built from) and, on the first
/// question about a state they hold no record for, have smplgpu_expand_state answer everything about that state
/// and its successors in one launch; the following calls are served from the record.  Same answers, callers
/// untouched; one launch per expansion.  The record is dropped whenever smplgpu_scene_epoch changes (robot, field,
/// walls, BFS run).
class ExpansionCache
{
public:
    /// deltas[n_prims][dof]: every primitive the action space may apply (converses included)
    ExpansionCache(smplgpu_ctx* ctx, int dof, const double* deltas, int n_prims, int cost_per_cell);
    bool ok() const { return m_ok; }
    void setCostPerCell(int c) { if (c != m_cost_per_cell) { m_cost_per_cell = c; m_n = 0; } }

    /// record of `q` if it is the current parent or one of its successors, else nullptr
    const smplgpu_succ_info* find(const double* q);
    /// find(q), or expand q as a new parent and return its own record (nullptr on a device error)
    const smplgpu_succ_info* get(const double* q);
    /// record of q AS A PARENT (the only record that carries isStateValid(q)), expanding q if need be
    const smplgpu_succ_info* parent(const double* q);
    /// record of the edge a -> b when b is a + a primitive (expanding a if need be), else nullptr
    const smplgpu_succ_info* edge(const double* a, const double* b);
    /// record whose planning-link position is exactly (x, y, z), else nullptr
    const smplgpu_succ_info* findLink(double x, double y, double z);

    long long launches() const { return m_launches; }
    long long hits() const { return m_hits; }

private:
    smplgpu_ctx* m_ctx;
    int m_dof, m_prims, m_cost_per_cell;
    bool m_ok = false;
    std::vector<double> m_deltas;
    const smplgpu_succ_info* m_info = nullptr;   // [m_n] records of the current parent, owned by the context
    int m_n = 0, m_hint = 0;
    std::vector<double> m_pending;   // the last state found as a SUCCESSOR entry: the likely next parent (see get())
    bool m_has_pending = false;
    bool isSuccessorOf(const double* a, const double* b) const;
    int64_t m_epoch = -1;
    long long m_launches = 0, m_hits = 0;
    bool valid();
    bool expand(const double* q);
};

class GpuCollisionSpace : public sbpl::motion::CollisionChecker, public sbpl::motion::CollisionDistanceExtension
{
public:
    GpuCollisionSpace(smplgpu_ctx* ctx, int dof) : m_ctx(ctx), m_dof(dof) { }

    // ---- sbpl::motion::CollisionChecker ----
    bool isStateValid(const sbpl::motion::RobotState& state, bool verbose = false) override;
    bool isStateValid(const sbpl::motion::RobotState& state, double& distToObst, bool verbose = false) override;
    bool isStateToStateValid(const sbpl::motion::RobotState& start, const sbpl::motion::RobotState& finish,
                             bool verbose = false) override;
    bool isStateToStateValid(const sbpl::motion::RobotState& a0, const sbpl::motion::RobotState& a1,
                             double& distToObst, int& distToObstCells, bool verbose = false) override;
    bool interpolatePath(const sbpl::motion::RobotState& start, const sbpl::motion::RobotState& finish,
                         std::vector<sbpl::motion::RobotState>& path) override;
    sbpl::motion::Extension* getExtension(size_t class_code) override;

    // ---- sbpl::motion::CollisionDistanceExtension (collision_checker.h:132-144) ----
    /// CollisionSpace::collisionDistance(state) (collision_space.cpp:496-500): the reference's clearance estimate
    double distanceToCollision(const sbpl::motion::RobotState& state) override;
    /// No class of the reference implements the motion form; here: the minimum of distanceToCollision over the
    /// waypoints isStateToStateValid checks (interpolatePath), all of them in one device call
    double distanceToCollision(const sbpl::motion::RobotState& start, const sbpl::motion::RobotState& finish) override;
    /// CollisionSpace::collisionDistance, as the reference names it
    double collisionDistance(const sbpl::motion::RobotState& state) { return distanceToCollision(state); }
    bool collisionDistances(const std::vector<sbpl::motion::RobotState>& states, std::vector<double>& dist);

    // ---- batched entry points: one GetSuccs submits every successor / edge at once ----
    bool isStatesValid(const std::vector<sbpl::motion::RobotState>& states, std::vector<uint8_t>& valid);
    bool isEdgesValid(const std::vector<sbpl::motion::RobotState>& starts,
                      const std::vector<sbpl::motion::RobotState>& finishes, std::vector<uint8_t>& valid);

    /// answer the per-call virtuals from one speculative launch per expansion (see ExpansionCache); not owned
    void setExpansionCache(ExpansionCache* cache) { m_cache = cache; }

    /// joint kinds of the planning variables, needed by interpolatePath (continuous => shortest arc)
    void setVariableInfo(const std::vector<int>& continuous, const std::vector<double>& motion_weights,
                         const std::vector<int>& var_types)
    { m_continuous = continuous; m_weights = motion_weights; m_types = var_types; }

private:
    smplgpu_ctx* m_ctx;
    int m_dof;
    std::vector<int> m_continuous, m_types;
    std::vector<double> m_weights;
    std::vector<double> m_buf0, m_buf1;
    ExpansionCache* m_cache = nullptr;
};

class GpuRobotModel : public sbpl::motion::ForwardKinematicsInterface
{
public:
    GpuRobotModel(smplgpu_ctx* ctx, RobotTables* tables, const std::string& planning_link);
    double minPosLimit(int jidx) const override { return m_min[jidx]; }
    double maxPosLimit(int jidx) const override { return m_max[jidx]; }
    bool hasPosLimit(int jidx) const override { return !m_cont[jidx]; }
    bool isContinuous(int jidx) const override { return m_cont[jidx] != 0; }
    double velLimit(int) const override { return 0.0; }
    double accLimit(int) const override { return 0.0; }
    bool checkJointLimits(const sbpl::motion::RobotState& state, bool verbose = false) override;
    /// only the planning link is supported (the one frame the hot path needs)
    bool computeFK(const sbpl::motion::RobotState& state, const std::string& name, std::vector<double>& pose) override;
    bool computePlanningLinkFK(const sbpl::motion::RobotState& state, std::vector<double>& pose) override;
    sbpl::motion::Extension* getExtension(size_t class_code) override;
    void setExpansionCache(ExpansionCache* cache) { m_cache = cache; }
private:
    smplgpu_ctx* m_ctx;
    ExpansionCache* m_cache = nullptr;
    std::string m_planning_link;
    std::vector<double> m_min, m_max;
    std::vector<int> m_cont;
};

class GpuBfsHeuristic : public sbpl::motion::RobotHeuristic
{
public:
    /// project: state id -> joint state of that lattice state (the role of PointProjectionExtension +
    /// ManipLattice::projectToPose, manip_lattice.cpp:1174-1206); returns false for unknown ids
    typedef std::function<bool(int, sbpl::motion::RobotState&)> StateLookup;

    GpuBfsHeuristic(smplgpu_ctx* ctx, const double origin[3], double res, const int dims[3]);
    bool init(StateLookup lookup, int goal_state_id);   // BfsHeuristic::init -> syncGridAndBfs
    void setInflationRadius(double r) { m_inflation_radius = r; }
    void setCostPerCell(int c) { m_cost_per_cell = c; if (m_cache) m_cache->setCostPerCell(c); }
    void setExpansionCache(ExpansionCache* cache) { m_cache = cache; if (cache) cache->setCostPerCell(m_cost_per_cell); }

    double getMetricStartDistance(double x, double y, double z) override;
    double getMetricGoalDistance(double x, double y, double z) override;
    void updateGoal(const sbpl::motion::GoalConstraint& goal) override;
    int GetGoalHeuristic(int state_id) override;
    int GetStartHeuristic(int) override { return 0; }
    int GetFromToHeuristic(int, int) override { return 0; }
    sbpl::motion::Extension* getExtension(size_t class_code) override;

    /// batched GetGoalHeuristic for a list of joint states
    bool goalHeuristics(const std::vector<sbpl::motion::RobotState>& states, std::vector<int>& h);
    int wallCount() const { return m_walls; }

private:
    smplgpu_ctx* m_ctx;
    double m_origin[3], m_res;
    int m_dims[3];
    double m_inflation_radius = 0.0;
    int m_cost_per_cell = 1;
    int m_walls = 0;
    StateLookup m_lookup;
    ExpansionCache* m_cache = nullptr;
    int m_goal_state_id = -1;
    double m_goal_xyz[3] = { 0, 0, 0 };
    void worldToGrid(double x, double y, double z, int cell[3]) const;
    int cellCost(const int cell[3]);
};

} // namespace smplhost

#endif
