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

class GpuCollisionSpace : public sbpl::motion::CollisionChecker
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

    // ---- batched entry points: one GetSuccs submits every successor / edge at once ----
    bool isStatesValid(const std::vector<sbpl::motion::RobotState>& states, std::vector<uint8_t>& valid);
    bool isEdgesValid(const std::vector<sbpl::motion::RobotState>& starts,
                      const std::vector<sbpl::motion::RobotState>& finishes, std::vector<uint8_t>& valid);

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
private:
    smplgpu_ctx* m_ctx;
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
    void setCostPerCell(int c) { m_cost_per_cell = c; }

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
    int m_goal_state_id = -1;
    double m_goal_xyz[3] = { 0, 0, 0 };
    void worldToGrid(double x, double y, double z, int cell[3]) const;
    int cellCost(const int cell[3]);
};

} // namespace smplhost

#endif
