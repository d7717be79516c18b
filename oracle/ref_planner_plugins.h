// ORACLE (test infrastructure only).  The action-space plug-in shared by ref_planner_shim.cpp and ref_dropin_shim.cpp:
// this fork's ManipLatticeActionSpace reads a motion-primitive format its own files do not have and rotates joints 0/1
// by joint 3 (SURVEY 8a defect 2); like oracle/lattice.cpp, ShimActionSpace follows the documented behaviour: every
// primitive is one waypoint parent + delta, weight 1 unless set, long primitives unless use_short_dist and
// RobotHeuristic::getMetricGoalDistance(planning link position) <= threshold (manip_lattice_action_space.cpp:376-449,
// 662-691), IK snap primitives off.
#ifndef ORACLE_REF_PLANNER_PLUGINS_H
#define ORACLE_REF_PLANNER_PLUGINS_H

#include <vector>

#include <smpl/graph/action_space.h>
#include <smpl/graph/robot_planning_space.h>
#include <smpl/heuristic/robot_heuristic.h>
#include <smpl/robot_model.h>
#include <smpl/search/lazy_arastar.h>

namespace {

using namespace sbpl::motion;

class ShimActionSpace : public ActionSpace
{
public:

    std::vector<std::vector<double>> deltas;   // file order, converse after each primitive (add_converse)
    std::vector<bool> is_short;
    std::vector<double> weight;                // per primitive (the fork's primitive files carry one, :182-190)
    bool use_short_dist = false;
    double short_dist_thresh = 0.0;
    ForwardKinematicsInterface* fk = nullptr;

    bool apply(const RobotState& parent, std::vector<Action>& actions) override
    {
        ActionsWeight w;
        return apply(parent, actions, w, -1);
    }

    bool apply(const RobotState& parent, std::vector<Action>& actions, ActionsWeight& weights, int) override
    {
        std::vector<double> pose;
        if (!fk->computePlanningLinkFK(parent, pose)) {
            return false;
        }
        // manip_lattice_action_space.cpp:385-396: distance of the planning link to the goal, from the first heuristic
        double goal_dist = 0.0;
        if (planningSpace()->numHeuristics() > 0) {
            goal_dist = planningSpace()->heuristic(0)->getMetricGoalDistance(pose[0], pose[1], pose[2]);
        }
        const bool near_goal = goal_dist <= short_dist_thresh;
        for (size_t p = 0; p < deltas.size(); ++p) {
            const bool active = is_short[p] ? (use_short_dist && near_goal) : !(use_short_dist && near_goal);
            if (!active) continue;
            Action action(1, parent);
            for (size_t j = 0; j < parent.size(); ++j) {
                action[0][j] = deltas[p][j] + parent[j];
            }
            actions.push_back(std::move(action));
            weights.push_back(p < weight.size() ? weight[p] : 1.0);
        }
        return true;
    }

    bool applyPredActions(const RobotState&, std::vector<Action>&, ActionsWeight&, int) override { return false; }
    void setMotionPlanRequestType(int) override { }
};

/// ManipLatticeActionSpace::addMotionPrim with add_converse (:201-228): the converse follows each primitive
inline std::vector<double>& ShimPrimWeights()
{
    static thread_local std::vector<double> w;   // set through ref*_set_prim_weights; empty = 1 each
    return w;
}

inline void FillPrimitives(ShimActionSpace& actions, const double* mprims, const uint8_t* short_flags, int n_prims, int dof)
{
    const std::vector<double>& w = ShimPrimWeights();
    for (int p = 0; p < n_prims; ++p) {
        std::vector<double> d(mprims + (size_t)p * dof, mprims + (size_t)(p + 1) * dof);
        const double wp = (int)w.size() == n_prims ? w[p] : 1.0;
        actions.deltas.push_back(d);
        actions.is_short.push_back(short_flags[p] != 0);
        actions.weight.push_back(wp);
        for (double& v : d) v = -v;
        actions.deltas.push_back(d);
        actions.is_short.push_back(short_flags[p] != 0);
        actions.weight.push_back(wp);
    }
}

/// The reference's in-tree lazy search (smpl/src/search/lazy_arastar.cpp) asks an sbpl::ILazySuccFun
/// (search/lazy_search_interface.h:54-65); no class of the reference implements it, so this is the adapter a user of
/// LazyARAStar writes: RobotPlanningSpace::GetLazySuccs / GetTrueCost (manip_lattice.cpp:1012-1167) as they are.
/// LazyARAStar has no expansion bound of its own: after `max_expansions` expansions this adapter returns no successors
/// and refuses evaluations, which drains OPEN (the same bound in the all-reference and the drop-in run).
struct LatticeLazySuccFun : public sbpl::ILazySuccFun
{
    RobotPlanningSpace* space = nullptr;
    int max_expansions = 0;
    int expansions = 0;
    int evaluations = 0;

    void GetLazySuccs(int state_id, std::vector<int>& succs, std::vector<int>& costs, std::vector<bool>& true_costs) override
    {
        if (expansions >= max_expansions) {
            return;
        }
        ++expansions;
        space->GetLazySuccs(state_id, &succs, &costs, &true_costs);
    }

    int GetSuccTrueCost(int state_id, int succ_id) override
    {
        if (expansions >= max_expansions) {
            return -1;
        }
        ++evaluations;
        return space->GetTrueCost(state_id, succ_id);
    }
};

/// one LazyARAStar query; returns true and fills solution / cost when the goal was reached
inline bool RunLazyARAStar(RobotPlanningSpace* space, RobotHeuristic* heur, double epsilon, int start_id, int goal_id,
                           int max_expansions, std::vector<int>& solution, int& cost, int& expansions, int& evaluations)
{
    LatticeLazySuccFun fun;
    fun.space = space;
    fun.max_expansions = max_expansions;
    sbpl::LazyARAStar search;
    search.eps_ = epsilon;
    bool found = false;
    if (sbpl::Init(search, &fun, heur)) {
        found = sbpl::Replan(search, start_id, goal_id, solution, cost) == 0;
    }
    expansions = fun.expansions;
    evaluations = fun.evaluations;
    for (auto* st : search.states_) {
        delete st;   // Clear() of lazy_arastar.cpp runs at the start of a query only
    }
    search.states_.clear();
    return found;
}

} // namespace

#endif
