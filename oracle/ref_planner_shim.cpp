// ORACLE (test infrastructure only -- never linked into the product path).
//
// One planning query through the REFERENCE's own planning stack, part of oracle/_ref/libref_collision.so:
//   ManipLattice + RobotPlanningSpace   smpl/src/graph/manip_lattice.cpp, robot_planning_space.cpp
//   BfsHeuristic + BFS_3D               smpl/src/heuristic/bfs_heuristic.cpp, robot_heuristic.cpp, smpl/src/bfs3d.cpp
//   ARAStar                             smpl/src/search/arastar.cpp
//   CollisionSpace                      the collision checker of ref_collision_shim.cpp
// all compiled where they lie (SBPL's base-class declarations: oracle/ref_stubs/sbpl).  It pins oracle/lattice.cpp
// and the BfsHeuristic of oracle/kdl_model.cpp: same path (state ids), cost, expansion count, lattice size and
// extracted joint path (tests/test_oracle_planner_reference.py, tests/golden/plans_reference.json).
//
// The RobotModel is the reference's own KDLRobotModel (sbpl_kdl_robot_model/src/kdl_robot_model.cpp, compiled where it
// lies against the KDL / kdl_parser stand-ins of oracle/ref_stubs/collision/kdl: KDL's arithmetic restated from its
// published sources, the reference's control flow its own).  One plug-in is NOT the reference's code, and says so:
//  * the ActionSpace (oracle/ref_planner_plugins.h).  This fork's ManipLatticeActionSpace reads a motion-primitive
//    format its own files do not have and rotates joints 0/1 by joint 3 (SURVEY 8a defect 2); like oracle/lattice.cpp,
//    ShimActionSpace follows the documented behaviour.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <smpl/debug/visualize.h>
#include <smpl/graph/manip_lattice.h>
#include <smpl/heuristic/bfs_heuristic.h>
#include <smpl/post_processing.h>
#include <smpl/search/arastar.h>

#include <sbpl_kdl_robot_model/kdl_robot_model.h>

#include "ref_planner_plugins.h"
#include "ref_collision_scene.h"

using namespace sbpl;
using namespace sbpl::motion;

// smpl/src/debug/visualize.cpp (marker publishing) is not compiled: visualisation is off
namespace sbpl {
namespace visual {
void visualize(Level, const Marker&) { }
void visualize(Level, const std::vector<Marker>&) { }
void visualize(Level, const visualization_msgs::Marker&) { }
void visualize(Level, const visualization_msgs::MarkerArray&) { }
void InitializeVizLocation(VizLocation* loc, const std::string&, Level level)
{
    loc->level = level;
    loc->enabled = false;
    loc->initialized = true;
    loc->handle = nullptr;
    loc->next = nullptr;
}
} // namespace visual
} // namespace sbpl

namespace {

/// the reference's own KDLRobotModel (sbpl_kdl_robot_model/src/kdl_robot_model.cpp over the KDL stand-in), set up as
/// call_planner.cpp does: init(urdf, planning joints, chain root, chain tip), kinematics -> planning transform,
/// planning link
bool SetupRobotModel(sbpl::motion::KDLRobotModel& robot, refcc_scene* s, const char* chain_root, const char* chain_tip,
                     const char* planning_link, const double* T_kin_to_planning /*3x4, nullable*/)
{
    if (!robot.init(s->robot_path, s->planning_joints, chain_root, chain_tip)) return false;
    if (T_kin_to_planning) {
        KDL::Frame f;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) f.M.data[3 * r + c] = T_kin_to_planning[4 * r + c];
            f.p.data[r] = T_kin_to_planning[4 * r + 3];
        }
        robot.setKinematicsToPlanningTransform(f, "planning");
    }
    return robot.setPlanningLink(planning_link);
}

} // namespace

// set by *_plan_lazy around a call of *_plan: the search is the reference's LazyARAStar over GetLazySuccs / GetTrueCost
static thread_local bool tl_lazy = false;
static thread_local int tl_lazy_evaluations = 0;

extern "C" {

/// same arguments and summary as oracle_plan (oracle/oracle_capi.cpp): out_summary = success, expansions, cost,
/// path length, lattice states created, extracted path length.  chain_* / T_kin_to_planning / xyz_offset as
/// oracle_scene_init_kdl.  Returns 0, or a negative step number when the reference refuses a step.
int refcc_plan(refcc_scene* s, const char* chain_root, const char* chain_tip, const char* planning_link,
               const double* T_kin_to_planning /*3x4*/, const double* xyz_offset,
               double inflation_radius, int cost_per_cell,
               const double* start, const double* goal_xyz,
               const double* resolutions, const double* mprims, const uint8_t* short_flags, int n_prims,
               int use_short_dist, double short_dist_thresh, double epsilon, int max_expansions,
               const double* xyz_tolerance, int32_t* out_summary, int32_t* path_ids, int max_path,
               double* path_states /* nullable: [max_path][dof] */)
{
    std::memset(out_summary, 0, 6 * sizeof(int32_t));
    const int dof = s->dof;

    KDLRobotModel robot;
    if (!SetupRobotModel(robot, s, chain_root, chain_tip, planning_link, T_kin_to_planning)) return -1;

    PlanningParams params;
    params.cost_per_cell = cost_per_cell;
    params.planning_link_sphere_radius = inflation_radius;

    ShimActionSpace actions;
    actions.fk = &robot;
    actions.use_short_dist = use_short_dist != 0;
    actions.short_dist_thresh = short_dist_thresh;
    FillPrimitives(actions, mprims, short_flags, n_prims, dof);

    ManipLattice space;
    const std::vector<double> res(resolutions, resolutions + dof);
    if (!space.init(&robot, s->cc.get(), &params, res, &actions)) return -3;
    if (!actions.init(&space)) return -4;

    BfsHeuristic heur;
    heur.setCostPerCell(cost_per_cell);
    heur.setInflationRadius(inflation_radius);
    if (!heur.init(&space, s->grid.get())) return -5;
    if (!space.insertHeuristic(&heur)) return -6;

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
    if (!space.setGoal(goal)) return -7;

    const RobotState st(start, start + dof);
    if (!space.setStart(st)) {
        out_summary[4] = (int)space.m_states.size();
        return 0;   // start outside the limits or in collision: no plan (PlannerInterface gives up the same way)
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
        if (!found) {
            return 0;
        }
    } else {
        ARAStar search(&space, &heur);
        search.set_initialsolution_eps(epsilon);
        if (search.set_start(space.getStartStateID()) == 0) return -8;
        if (search.set_goal(space.getGoalStateID()) == 0) return -9;
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

/// refcc_plan with the reference's lazy successors: ManipLattice::GetLazySuccs / GetTrueCost (manip_lattice.cpp:1012-1167)
/// under the in-tree LazyARAStar (search/lazy_arastar.cpp, weight `epsilon`, one pass).  out_summary[1] = expansions
/// (GetLazySuccs calls); refcc_last_lazy_evaluations() = GetTrueCost calls of the last query on this thread.
int refcc_plan_lazy(refcc_scene* s, const char* chain_root, const char* chain_tip, const char* planning_link,
                    const double* T_kin_to_planning, const double* xyz_offset,
                    double inflation_radius, int cost_per_cell,
                    const double* start, const double* goal_xyz,
                    const double* resolutions, const double* mprims, const uint8_t* short_flags, int n_prims,
                    int use_short_dist, double short_dist_thresh, double epsilon, int max_expansions,
                    const double* xyz_tolerance, int32_t* out_summary, int32_t* path_ids, int max_path,
                    double* path_states)
{
    tl_lazy = true;
    const int r = refcc_plan(s, chain_root, chain_tip, planning_link, T_kin_to_planning, xyz_offset, inflation_radius,
                             cost_per_cell, start, goal_xyz, resolutions, mprims, short_flags, n_prims, use_short_dist,
                             short_dist_thresh, epsilon, max_expansions, xyz_tolerance, out_summary, path_ids, max_path,
                             path_states);
    tl_lazy = false;
    return r;
}

/// action weights of the primitives for the following refcc_plan calls on this thread (n = 0 clears them)
void refcc_set_prim_weights(const double* weights, int n)
{
    ShimPrimWeights().assign(weights, weights + (n > 0 ? n : 0));
}

int refcc_last_lazy_evaluations(void)
{
    return tl_lazy_evaluations;
}

/// BfsHeuristic::GetGoalHeuristic (bfs_heuristic.cpp:148-163, 355-366) of the reference for n joint states: each state
/// becomes a lattice state of the reference's ManipLattice (getOrCreateState), the goal is set through
/// ManipLattice::setGoal -> BfsHeuristic::updateGoal -> BFS_3D::run, and the heuristic is asked by state id; h[n] is
/// followed by the value of the goal state itself (h[n]).  Also returns getMetricGoalDistance of every state's planning
/// frame position in metric[n].
int refcc_goal_heuristics(refcc_scene* s, const char* chain_root, const char* chain_tip, const char* planning_link,
                          const double* T_kin_to_planning, const double* xyz_offset, double inflation_radius,
                          int cost_per_cell, const double* goal_xyz, const double* resolutions,
                          const double* q, int n, int32_t* h, double* metric)
{
    const int dof = s->dof;
    KDLRobotModel robot;
    if (!SetupRobotModel(robot, s, chain_root, chain_tip, planning_link, T_kin_to_planning)) return -1;
    PlanningParams params;
    params.cost_per_cell = cost_per_cell;
    ShimActionSpace actions;
    actions.fk = &robot;
    ManipLattice space;
    const std::vector<double> res(resolutions, resolutions + dof);
    if (!space.init(&robot, s->cc.get(), &params, res, &actions)) return -3;
    if (!actions.init(&space)) return -4;
    BfsHeuristic heur;
    heur.setCostPerCell(cost_per_cell);
    heur.setInflationRadius(inflation_radius);
    if (!heur.init(&space, s->grid.get())) return -5;
    if (!space.insertHeuristic(&heur)) return -6;
    GoalConstraint goal;
    goal.type = GoalType::XYZ_GOAL;
    goal.pose.assign(6, 0.0);
    goal.tgt_off_pose.assign(6, 0.0);
    for (int i = 0; i < 3; ++i) {
        goal.pose[i] = goal_xyz[i];
        goal.tgt_off_pose[i] = goal_xyz[i];
        goal.xyz_offset[i] = xyz_offset[i];
        goal.xyz_tolerance[i] = 0.015;
        goal.rpy_tolerance[i] = 0.0;
    }
    if (!space.setGoal(goal)) return -7;
    RobotCoord coord(dof);
    std::vector<double> pose;
    for (int i = 0; i < n; ++i) {
        const RobotState st(q + (size_t)i * dof, q + (size_t)(i + 1) * dof);
        space.stateToCoord(st, coord);
        const int id = space.createHashEntry(coord, st);   // one lattice state per input state, whatever its cell
        h[i] = heur.GetGoalHeuristic(id);
        if (!space.computePlanningFrameFK(st, pose)) return -8;
        metric[i] = heur.getMetricGoalDistance(pose[0], pose[1], pose[2]);
    }
    h[n] = heur.GetGoalHeuristic(space.getGoalStateID());
    return 0;
}

/// KDLRobotModel::computePlanningLinkFK (kdl_robot_model.cpp:400-423) and checkJointLimits (:326-337) for n states:
/// pose6 [n][6] = x y z roll pitch yaw, within [n]
int refcc_kdl_fk_and_limits(refcc_scene* s, const char* chain_root, const char* chain_tip, const char* planning_link,
                            const double* T_kin_to_planning, const double* q, int n, double* pose6, uint8_t* within)
{
    KDLRobotModel robot;
    if (!SetupRobotModel(robot, s, chain_root, chain_tip, planning_link, T_kin_to_planning)) return -1;
    std::vector<double> pose;
    for (int i = 0; i < n; ++i) {
        const RobotState st(q + (size_t)i * s->dof, q + (size_t)(i + 1) * s->dof);
        if (!robot.computePlanningLinkFK(st, pose)) return -2;
        for (int k = 0; k < 6; ++k) pose6[(size_t)i * 6 + k] = pose[k];
        within[i] = robot.checkJointLimits(st) ? 1 : 0;
    }
    return 0;
}

/// limits as the planner sees them (KDLRobotModel::minPosLimit / maxPosLimit / isContinuous)
int refcc_kdl_limits(refcc_scene* s, const char* chain_root, const char* chain_tip, double* lo, double* hi, uint8_t* continuous)
{
    KDLRobotModel robot;
    if (!robot.init(s->robot_path, s->planning_joints, chain_root, chain_tip)) return -1;
    for (int i = 0; i < s->dof; ++i) {
        lo[i] = robot.minPosLimit(i);
        hi[i] = robot.maxPosLimit(i);
        continuous[i] = robot.isContinuous(i) ? 1 : 0;
    }
    return 0;
}

/// ShortcutPath(rm, cc, pin, pout, type) / InterpolatePath(cc, path) of smpl/src/post_processing.cpp on a joint-space
/// path (n x dof): kind 0 = JOINT_SPACE, 1 = JOINT_POSITION_VELOCITY_SPACE, 2 = InterpolatePath.  out: [max_out][dof];
/// returns the number of points (negative: too many for max_out, or the reference reported failure).
int refcc_post_process(refcc_scene* s, const char* chain_root, const char* chain_tip, const char* planning_link,
                       const double* path, int n, int kind, double* out, int max_out)
{
    const int dof = s->dof;
    KDLRobotModel robot;
    if (!SetupRobotModel(robot, s, chain_root, chain_tip, planning_link, nullptr)) return -1;
    std::vector<RobotState> pin(n), pout;
    for (int i = 0; i < n; ++i) pin[i].assign(path + (size_t)i * dof, path + (size_t)(i + 1) * dof);
    if (kind == 2) {
        pout = pin;
        if (!InterpolatePath(*s->cc, pout)) return -3;
    } else {
        ShortcutPath(&robot, s->cc.get(), pin, pout,
                     kind == 0 ? ShortcutType::JOINT_SPACE : ShortcutType::JOINT_POSITION_VELOCITY_SPACE);
    }
    if ((int)pout.size() > max_out) return -(int)pout.size();
    for (size_t i = 0; i < pout.size(); ++i) {
        for (int d = 0; d < dof; ++d) out[i * dof + d] = pout[i][d];
    }
    return (int)pout.size();
}

} // extern "C"
