// ORACLE (test infrastructure only -- never linked into the product path).
//
// Minimal CPU restatement of the CALLER of the hot path, used to check
// query-level parity (identical paths, costs, expansion counts) and to time
// "plan queries/s" on the host:
//   ManipLattice      smpl/src/graph/manip_lattice.cpp:72-150 (init), 219-313 (GetSuccs),
//                     1245-1289 (coord <-> state), 1511-1580 (checkAction), 1582-1696 (isGoal),
//                     1944-1980 (setStart), 2018-2160 (extractPath)
//   ManipLatticeActionSpace  smpl/src/graph/manip_lattice_action_space.cpp:201-228 (addMotionPrim),
//                     376-449 (apply), 507-573 (getAction), 662-691 (mprimActive)
//   ARAStar           oracle/arastar.h (pinned against the reference's own arastar.cpp)
//   lazy successors   manip_lattice.cpp:1012-1090 (GetLazySuccs), 1094-1167 (GetTrueCost); search: oracle/lazy_arastar.h
//
// Decisions (fork defect 2): motion primitives use the documented plain format
// (delta per joint, a weight -- 1 unless given --, no base rotation hack, converse added after each primitive); the IK "snap"
// primitives need third-party IK and are outside the parity set; the goal is an XYZ_GOAL (position within
// xyz_tolerance of the target offset pose, manip_lattice.cpp:1673-1687); the search stops at the first
// solution of the initial epsilon (improve = false) or after max_expansions, so results do not depend on
// wall-clock time.  parity unpinned (the reference records no expected plans).
#ifndef ORACLE_LATTICE_H
#define ORACLE_LATTICE_H

#include <map>
#include <vector>

#include "arastar.h"
#include "lazy_arastar.h"
#include "collision_space.h"
#include "kdl_model.h"

namespace oracle {

struct MotionPrim
{
    std::vector<double> delta;
    bool short_dist;
    double weight;   // the primitive file's weight column: an edge costs (int)(1000 * weight), manip_lattice.cpp:1414-1437
    MotionPrim() : short_dist(false), weight(1.0) { }
};

struct PlanParams
{
    std::vector<double> resolutions;   // per planning variable (radians / metres)
    std::vector<MotionPrim> mprims;    // in file order; converses are added by the lattice
    bool use_short_dist;
    double short_dist_thresh;
    double epsilon;
    int max_expansions;
    double xyz_tolerance[3];
};

struct PlanResult
{
    bool success;
    int expansions;
    int cost;
    std::vector<int> path_ids;
    std::vector<std::vector<double>> path_states;
    int num_states;
    int evaluations;   // lazy search: GetTrueCost calls
    PlanResult() : success(false), expansions(0), cost(0), num_states(0), evaluations(0) { }
};

class ManipLatticePlanner
{
public:
    ManipLatticePlanner(CollisionSpace* cc, KDLRobotModel* robot, BfsHeuristic* heur,
                        const double xyz_offset[3], int cost_per_cell, const PlanParams& params);

    /// PlannerInterface::planToPose reduced to: setGoal (BFS from the goal cell), setStart, ARA* replan
    PlanResult plan(const std::vector<double>& start, const double goal_xyz[3]);

    /// the same query through the lazy successors (ManipLattice::GetLazySuccs / GetTrueCost, manip_lattice.cpp:1012-1167)
    /// under the reference's in-tree LazyARAStar (oracle/lazy_arastar.h), bounded by max_expansions in the successor
    /// function exactly as oracle/ref_planner_plugins.h:LatticeLazySuccFun bounds the reference's run
    PlanResult planLazy(const std::vector<double>& start, const double goal_xyz[3]);

private:
    struct LatticeState { std::vector<int> coord; std::vector<double> state; };
    CollisionSpace* m_cc;
    KDLRobotModel* m_robot;
    BfsHeuristic* m_heur;
    double m_xyz_offset[3];
    PlanParams m_params;
    std::vector<std::vector<double>> m_prim_deltas;
    std::vector<bool> m_prim_short;
    std::vector<int> m_prim_cost;      // GetSuccs' edge cost per primitive, converses included

    std::vector<double> m_min_limits, m_max_limits, m_coord_deltas;
    std::vector<bool> m_continuous, m_bounded;
    std::vector<int> m_coord_vals;

    std::vector<LatticeState> m_states;
    std::map<std::vector<int>, int> m_coord_to_id;
    int m_goal_state_id, m_start_state_id;
    double m_goal[3];

    void stateToCoord(const std::vector<double>& state, std::vector<int>& coord) const;
    int getOrCreateState(const std::vector<int>& coord, const std::vector<double>& state);
    bool isGoal(const std::vector<double>& state) const;
    void getSuccs(int state_id, std::vector<int>& succs, std::vector<int>& costs);
    bool primActive(size_t p, bool near_goal) const;
    void getLazySuccs(int state_id, std::vector<int>& succs, std::vector<int>& costs, std::vector<bool>& true_costs);
    int getTrueCost(int parent_id, int child_id);
    bool begin(const std::vector<double>& start, const double goal_xyz[3], PlanResult& res);
    int goalHeuristic(int state_id) const;
    bool extractPath(const std::vector<int>& ids, std::vector<std::vector<double>>& path) const;

};

} // namespace oracle

#endif
