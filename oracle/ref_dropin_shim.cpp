// ORACLE (test infrastructure only -- never linked into the product path).
//
// The drop-in claim, run literally: the REFERENCE's own ManipLattice + RobotPlanningSpace + ARAStar (compiled from
// /root/reference, oracle/_ref/libref_collision.so) plan a query with the PRODUCT's plug-ins behind the reference's own
// interfaces: smplhost::GpuCollisionSpace as sbpl::motion::CollisionChecker, GpuRobotModel as RobotModel +
// ForwardKinematicsInterface, GpuBfsHeuristic as RobotHeuristic (smpl_b200/host/gpu_adapters.cpp, compiled here
// against the reference's REAL headers with -DSMPLHOST_REFERENCE_HEADERS instead of the restated
// smpl_b200/host/smpl/interfaces.h) over libsmplgpu.so's C ABI.  Every isStateValid / isStateToStateValid /
// GetGoalHeuristic / computePlanningLinkFK / checkJointLimits call of the reference's search lands in a CUDA kernel.
// tests/test_gpu_dropin.py requires the same success flag, expansion count, cost, lattice size, id path and joint path as
// the all-reference run of ref_planner_shim.cpp (tests/golden/plans_reference.json).  -> oracle/_ref/libref_dropin.so
//
// The one addition to the adapters: ManipLattice::init insists on an InverseKinematicsInterface
// (manip_lattice.cpp:100-104), which the hot path never calls; DropInRobotModel adds one that always fails.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <smpl/graph/manip_lattice.h>
#include <smpl/search/arastar.h>

#include "../smpl_b200/host/gpu_adapters.h"
#include "ref_planner_plugins.h"

using namespace sbpl;
using namespace sbpl::motion;

namespace {

class DropInRobotModel : public smplhost::GpuRobotModel, public InverseKinematicsInterface
{
public:
    DropInRobotModel(smplgpu_ctx* ctx, smplhost::RobotTables* tables, const std::string& planning_link) :
        smplhost::GpuRobotModel(ctx, tables, planning_link) { }
    // RobotModel is a virtual base of both interfaces: the adapter's implementations are the final overriders
    double minPosLimit(int j) const override { return smplhost::GpuRobotModel::minPosLimit(j); }
    double maxPosLimit(int j) const override { return smplhost::GpuRobotModel::maxPosLimit(j); }
    bool hasPosLimit(int j) const override { return smplhost::GpuRobotModel::hasPosLimit(j); }
    bool isContinuous(int j) const override { return smplhost::GpuRobotModel::isContinuous(j); }
    double velLimit(int j) const override { return smplhost::GpuRobotModel::velLimit(j); }
    double accLimit(int j) const override { return smplhost::GpuRobotModel::accLimit(j); }
    bool checkJointLimits(const RobotState& s, bool v = false) override { return smplhost::GpuRobotModel::checkJointLimits(s, v); }
    bool computeIK(const std::vector<double>&, const RobotState&, RobotState&, ik_option::IkOption) override { return false; }
    bool computeIK(const std::vector<double>&, const RobotState&, std::vector<RobotState>&, ik_option::IkOption) override { return false; }
    Extension* getExtension(size_t class_code) override
    {
        if (class_code == GetClassCode<InverseKinematicsInterface>()) return static_cast<InverseKinematicsInterface*>(this);
        return smplhost::GpuRobotModel::getExtension(class_code);
    }
};

/// INTEGRATION.md section 4, as a subclass instead of an edit: the reference's ManipLattice with GetSuccs collecting
/// every single-waypoint action first and checking them in ONE GpuCollisionSpace::isEdgesValid call (one launch per
/// expansion instead of one per successor).  Everything else -- stateToCoord, getOrCreateState, isGoal, cost, the
/// order successors are emitted in -- is the reference's own code (manip_lattice.cpp:219-313, 1511-1580), reached
/// through -fno-access-control.
class BatchedManipLattice : public ManipLattice
{
public:
    smplhost::GpuCollisionSpace* gpu = nullptr;
    long long batched_calls = 0, edges_submitted = 0;

    void GetSuccs(int state_id, std::vector<int>* succs, std::vector<int>* costs) override
    {
        if (!gpu) {
            ManipLattice::GetSuccs(state_id, succs, costs);   // the reference's own loop: one call per successor
            return;
        }
        if (state_id == m_goal_state_id) {
            return;   // goal state is absorbing
        }
        ManipLatticeState* parent_entry = m_states[state_id];
        std::vector<Action> actions;
        ActionsWeight weights;
        if (!m_actions->apply(parent_entry->state, actions, weights, -1)) {
            return;
        }
        // checkAction, first half: joint limits of every waypoint; multi-waypoint actions (none of the motion
        // primitives) would keep the per-action path
        std::vector<size_t> kept;
        std::vector<RobotState> starts, finishes;
        for (size_t i = 0; i < actions.size(); ++i) {
            if (actions[i].size() != 1 || !robot()->checkJointLimits(actions[i][0])) {
                continue;
            }
            kept.push_back(i);
            starts.push_back(parent_entry->state);
            finishes.push_back(actions[i][0]);
        }
        // checkAction, second half: the edges parent -> waypoint, all at once
        std::vector<uint8_t> ok;
        if (!starts.empty() && !gpu->isEdgesValid(starts, finishes, ok)) {
            return;
        }
        ++batched_calls;
        edges_submitted += (long long)starts.size();
        RobotCoord succ_coord(robot()->jointVariableCount(), 0);
        for (size_t k = 0; k < kept.size(); ++k) {
            if (!ok[k]) {
                continue;
            }
            const Action& action = actions[kept[k]];
            stateToCoord(action.back(), succ_coord);
            const int succ_state_id = getOrCreateState(succ_coord, action.back());
            ManipLatticeState* succ_entry = getHashEntry(succ_state_id);
            const bool is_goal_succ = isGoal(action.back());
            succs->push_back(is_goal_succ ? m_goal_state_id : succ_state_id);
            costs->push_back(cost(parent_entry, succ_entry, weights[kept[k]], is_goal_succ));
        }
    }
};

std::vector<std::string> SplitCsv(const char* s)
{
    std::vector<std::string> out;
    std::string item;
    for (const char* p = s ? s : ""; ; ++p) {
        if (*p == ',' || *p == 0) {
            if (!item.empty()) out.push_back(item);
            item.clear();
            if (*p == 0) break;
        } else {
            item.push_back(*p);
        }
    }
    return out;
}

} // namespace

// set by *_plan_lazy around a call of *_plan: the search is the reference's LazyARAStar over GetLazySuccs / GetTrueCost
static thread_local bool tl_lazy = false;
static thread_local int tl_lazy_evaluations = 0;

extern "C" {

/// batched_get_succs == 1 runs the reference's lattice with the INTEGRATION.md section-4 edit (BatchedManipLattice);
/// out_summary[6], [7] then hold the number of batched calls and of edges submitted.
/// batched_get_succs == 2 runs the UNCHANGED reference lattice (one virtual call per question) with the adapters
/// sharing a smplhost::ExpansionCache built from the action space's primitive table: one speculative launch per
/// expansion answers all of its questions; out_summary[6], [7] = cache launches, cache hits.
/// ctx: a smplgpu context that already holds the robot tables, the distance field and the planning chain (set up
/// through the ABI by the caller).  Arguments and summary as refcc_plan / oracle_plan.  Returns 0, or a negative
/// step number when a step is refused.
int refdrop_plan(smplgpu_ctx* ctx, const char* robot_path, const char* group, const char* planning_joints_csv,
                 const char* planning_link, const double* grid_origin, double grid_res, const int32_t* grid_dims,
                 double inflation_radius, int cost_per_cell,
                 const double* start, const double* goal_xyz, const double* xyz_offset,
                 const double* resolutions, const double* mprims, const uint8_t* short_flags, int n_prims,
                 int use_short_dist, double short_dist_thresh, double epsilon, int max_expansions,
                 const double* xyz_tolerance, int32_t* out_summary, int32_t* path_ids, int max_path,
                 double* path_states /* nullable */, int batched_get_succs)
{
    std::memset(out_summary, 0, 8 * sizeof(int32_t));
    const std::vector<std::string> joints = SplitCsv(planning_joints_csv);
    const int dof = (int)joints.size();

    // limits of the planning variables for the RobotModel adapter
    smplhost::RobotTables tables;
    std::string err;
    if (!tables.load(robot_path, &err)) return -1;
    if (!tables.configure(group, joints, &err)) return -2;

    DropInRobotModel robot(ctx, &tables, planning_link);
    robot.setPlanningJoints(joints);
    smplhost::GpuCollisionSpace checker(ctx, dof);

    PlanningParams params;
    params.cost_per_cell = cost_per_cell;
    params.planning_link_sphere_radius = inflation_radius;

    ShimActionSpace actions;
    actions.fk = &robot;
    actions.use_short_dist = use_short_dist != 0;
    actions.short_dist_thresh = short_dist_thresh;
    FillPrimitives(actions, mprims, short_flags, n_prims, dof);

    // batched_get_succs: the reference's lattice with the section-4 edit (BatchedManipLattice) instead of the plain one
    // mode 2: the adapters answer from one smplgpu_expand_state record per expansion
    std::vector<double> flat;
    for (const auto& d : actions.deltas) flat.insert(flat.end(), d.begin(), d.end());
    std::unique_ptr<smplhost::ExpansionCache> cache;
    if (batched_get_succs == 2) {
        cache.reset(new smplhost::ExpansionCache(ctx, dof, flat.data(), (int)actions.deltas.size(), cost_per_cell));
        if (!cache->ok()) return -11;
        robot.setExpansionCache(cache.get());
        checker.setExpansionCache(cache.get());
    }

    BatchedManipLattice space;
    space.gpu = batched_get_succs == 1 ? &checker : nullptr;
    const std::vector<double> res(resolutions, resolutions + dof);
    if (!space.init(&robot, &checker, &params, res, &actions)) return -3;
    if (!actions.init(&space)) return -4;

    const int dims[3] = { grid_dims[0], grid_dims[1], grid_dims[2] };
    smplhost::GpuBfsHeuristic heur(ctx, grid_origin, grid_res, dims);
    heur.setCostPerCell(cost_per_cell);
    heur.setInflationRadius(inflation_radius);
    if (batched_get_succs == 2) heur.setExpansionCache(cache.get());
    if (!heur.RobotHeuristic::init(&space)) return -5;
    if (!heur.init([&space](int id, RobotState& q) {
            if (id < 0 || id >= (int)space.m_states.size() || !space.m_states[id]) return false;
            q = space.extractState(id);
            return !q.empty();
        }, space.getGoalStateID())) return -6;
    if (!space.insertHeuristic(&heur)) return -7;

    GoalConstraint goal;
    goal.type = GoalType::XYZ_GOAL;
    goal.pose.assign(6, 0.0);
    goal.tgt_off_pose.assign(6, 0.0);
    for (int i = 0; i < 3; ++i) {
        goal.pose[i] = goal_xyz[i];
        goal.tgt_off_pose[i] = goal_xyz[i];
        goal.xyz_offset[i] = xyz_offset[i];
        goal.xyz_tolerance[i] = xyz_tolerance[i];
        goal.rpy_tolerance[i] = 0.0;
    }
    if (!space.setGoal(goal)) return -8;

    const RobotState st(start, start + dof);
    if (!space.setStart(st)) {
        out_summary[4] = (int)space.m_states.size();
        return 0;
    }

    std::vector<int> solution;
    int solcost = 0;
    if (tl_lazy) {
        // the reference's lazy successors (GetLazySuccs / GetTrueCost) under its in-tree LazyARAStar
        int expansions = 0;
        tl_lazy_evaluations = 0;
        const bool found = RunLazyARAStar(&space, &heur, epsilon, space.getStartStateID(), space.getGoalStateID(),
                                          max_expansions, solution, solcost, expansions, tl_lazy_evaluations);
        out_summary[1] = expansions;
        out_summary[4] = (int)space.m_states.size();
        out_summary[6] = cache ? (int)cache->launches() : 0;
        out_summary[7] = cache ? (int)cache->hits() : 0;
        if (!found) {
            return 0;
        }
    } else {
        ARAStar search(&space, &heur);
        search.set_initialsolution_eps(epsilon);
        if (search.set_start(space.getStartStateID()) == 0) return -9;
        if (search.set_goal(space.getGoalStateID()) == 0) return -10;
        ARAStar::TimeParameters tp;
        tp.bounded = true;
        tp.improve = false;
        tp.type = ARAStar::TimeParameters::EXPANSIONS;
        tp.max_expansions_init = max_expansions;
        tp.max_expansions = max_expansions;
        tp.max_allowed_time_init = sbpl::clock::duration::zero();
        tp.max_allowed_time = sbpl::clock::duration::zero();
        const int ret = search.replan(tp, &solution, &solcost);
        out_summary[1] = search.get_n_expands();
        out_summary[4] = (int)space.m_states.size();
        out_summary[6] = cache ? (int)cache->launches() : (int)space.batched_calls;
        out_summary[7] = cache ? (int)cache->hits() : (int)space.edges_submitted;
        if (!ret || solcost >= INFINITECOST) {
            return 0;
        }
    }
    out_summary[0] = 1;
    out_summary[2] = solcost;
    out_summary[3] = (int)solution.size();
    for (int i = 0; i < (int)solution.size() && i < max_path; ++i) {
        path_ids[i] = solution[i];
    }
    if (path_states) {
        std::vector<RobotState> path;
        if (space.extractPath(solution, path)) {
            out_summary[5] = (int)path.size();
            for (int i = 0; i < (int)path.size() && i < max_path; ++i) {
                for (int d = 0; d < dof; ++d) path_states[(size_t)i * dof + d] = path[i][d];
            }
        }
    }
    return 0;
}

/// refdrop_plan with the reference's lazy successors (GetLazySuccs / GetTrueCost) under its in-tree LazyARAStar, the
/// lattice and the search untouched, every question answered by the adapters (see refcc_plan_lazy).
int refdrop_plan_lazy(smplgpu_ctx* ctx, const char* robot_path, const char* group, const char* planning_joints_csv,
                      const char* planning_link, const double* grid_origin, double grid_res, const int32_t* grid_dims,
                      double inflation_radius, int cost_per_cell,
                      const double* start, const double* goal_xyz, const double* xyz_offset,
                      const double* resolutions, const double* mprims, const uint8_t* short_flags, int n_prims,
                      int use_short_dist, double short_dist_thresh, double epsilon, int max_expansions,
                      const double* xyz_tolerance, int32_t* out_summary, int32_t* path_ids, int max_path,
                      double* path_states, int batched_get_succs)
{
    tl_lazy = true;
    const int r = refdrop_plan(ctx, robot_path, group, planning_joints_csv, planning_link, grid_origin, grid_res, grid_dims,
                               inflation_radius, cost_per_cell, start, goal_xyz, xyz_offset, resolutions, mprims, short_flags,
                               n_prims, use_short_dist, short_dist_thresh, epsilon, max_expansions, xyz_tolerance, out_summary,
                               path_ids, max_path, path_states, batched_get_succs);
    tl_lazy = false;
    return r;
}

int refdrop_last_lazy_evaluations(void)
{
    return tl_lazy_evaluations;
}

} // extern "C"
