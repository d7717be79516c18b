// ORACLE (test infrastructure only -- never linked into the product path).
#include "lattice.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace oracle {


ManipLatticePlanner::ManipLatticePlanner(
    CollisionSpace* cc, KDLRobotModel* robot, BfsHeuristic* heur,
    const double xyz_offset[3], int /*cost_per_cell*/, const PlanParams& params)
:
    m_cc(cc), m_robot(robot), m_heur(heur), m_params(params),
    m_goal_state_id(-1), m_start_state_id(-1)
{
    for (int i = 0; i < 3; ++i) m_xyz_offset[i] = xyz_offset[i];
    // addMotionPrim(..., add_converse = true): each primitive is followed by its negation
    for (const MotionPrim& p : params.mprims) {
        m_prim_deltas.push_back(p.delta);
        m_prim_short.push_back(p.short_dist);
        m_prim_cost.push_back((int)(1000 * p.weight));   // DefaultCostMultiplier * actionWeight, truncated
        std::vector<double> neg(p.delta);
        for (double& v : neg) v *= -1.0;
        m_prim_deltas.push_back(neg);
        m_prim_short.push_back(p.short_dist);
        m_prim_cost.push_back((int)(1000 * p.weight));
    }
    // ManipLattice::init (manip_lattice.cpp:105-146)
    const size_t n = robot->min_limits.size();
    m_min_limits = robot->min_limits;
    m_max_limits = robot->max_limits;
    m_continuous.resize(n);
    m_bounded.resize(n);
    m_coord_vals.resize(n);
    m_coord_deltas.resize(n);
    for (size_t v = 0; v < n; ++v) {
        m_continuous[v] = robot->continuous[v];
        m_bounded[v] = !robot->continuous[v]; // KDLRobotModel::hasPosLimit
        const double res = params.resolutions[v];
        if (m_continuous[v]) {
            m_coord_vals[v] = (int)std::round((2.0 * M_PI) / res);
            m_coord_deltas[v] = (2.0 * M_PI) / (double)m_coord_vals[v];
        } else if (m_bounded[v]) {
            const double span = std::fabs(m_max_limits[v] - m_min_limits[v]);
            m_coord_vals[v] = std::max(1, (int)std::round(span / res));
            m_coord_deltas[v] = span / (double)m_coord_vals[v];
        } else {
            m_coord_vals[v] = std::numeric_limits<int>::max();
            m_coord_deltas[v] = res;
        }
    }
}

/// manip_lattice.cpp:1263-1289
void ManipLatticePlanner::stateToCoord(const std::vector<double>& state, std::vector<int>& coord) const
{
    coord.resize(state.size());
    for (size_t i = 0; i < state.size(); ++i) {
        if (m_continuous[i]) {
            double pos_angle = normalize_angle(state[i]);
            if (pos_angle < 0.0) {
                pos_angle += 2.0 * M_PI; // angles::normalize_angle_positive
            }
            coord[i] = (int)((pos_angle + m_coord_deltas[i] * 0.5) / m_coord_deltas[i]);
            if (coord[i] == m_coord_vals[i]) {
                coord[i] = 0;
            }
        } else if (!m_bounded[i]) {
            if (state[i] >= 0.0) {
                coord[i] = (int)(state[i] / m_coord_deltas[i] + 0.5);
            } else {
                coord[i] = (int)(state[i] / m_coord_deltas[i] - 0.5);
            }
        } else {
            coord[i] = (int)(((state[i] - m_min_limits[i]) / m_coord_deltas[i]) + 0.5);
        }
    }
}

int ManipLatticePlanner::getOrCreateState(const std::vector<int>& coord, const std::vector<double>& state)
{
    auto it = m_coord_to_id.find(coord);
    if (it != m_coord_to_id.end()) {
        return it->second;
    }
    const int id = (int)m_states.size();
    m_states.push_back(LatticeState{ coord, state });
    m_coord_to_id[coord] = id;
    return id;
}

/// manip_lattice.cpp:1673-1687 (XYZ_GOAL)
bool ManipLatticePlanner::isGoal(const std::vector<double>& state) const
{
    std::vector<double> pose;
    m_robot->computePlanningLinkFK(state, pose);
    pose = GetTargetOffsetPose(pose, m_xyz_offset);
    return std::fabs(pose[0] - m_goal[0]) <= m_params.xyz_tolerance[0] &&
           std::fabs(pose[1] - m_goal[1]) <= m_params.xyz_tolerance[1] &&
           std::fabs(pose[2] - m_goal[2]) <= m_params.xyz_tolerance[2];
}

/// BfsHeuristic::GetGoalHeuristic over ManipLattice::projectToPose (bfs_heuristic.cpp:148-163,
/// manip_lattice.cpp:1174-1206): the goal state projects to the goal pose itself
int ManipLatticePlanner::goalHeuristic(int state_id) const
{
    if (state_id == m_goal_state_id) {
        return m_heur->getGoalHeuristicAt(m_goal[0], m_goal[1], m_goal[2]);
    }
    std::vector<double> pose;
    m_robot->computePlanningLinkFK(m_states[state_id].state, pose);
    pose = GetTargetOffsetPose(pose, m_xyz_offset);
    return m_heur->getGoalHeuristicAt(pose[0], pose[1], pose[2]);
}

/// ManipLattice::GetSuccs + ManipLatticeActionSpace::apply + ManipLattice::checkAction
void ManipLatticePlanner::getSuccs(int state_id, std::vector<int>& succs, std::vector<int>& costs)
{
    if (state_id == m_goal_state_id) {
        return; // goal state is absorbing
    }
    const std::vector<double> parent = m_states[state_id].state;

    std::vector<double> pose;
    m_robot->computePlanningLinkFK(parent, pose);
    const double goal_dist = m_heur->getMetricGoalDistance(pose[0], pose[1], pose[2]);
    const bool near_goal = goal_dist <= m_params.short_dist_thresh;

    std::vector<double> succ(parent.size());
    std::vector<int> coord;
    for (size_t p = 0; p < m_prim_deltas.size(); ++p) {
        // mprimActive (manip_lattice_action_space.cpp:662-691)
        bool active;
        if (!m_prim_short[p]) {
            active = !(m_params.use_short_dist && near_goal);
        } else {
            active = m_params.use_short_dist && near_goal;
        }
        if (!active) {
            continue;
        }
        for (size_t j = 0; j < parent.size(); ++j) {
            succ[j] = m_prim_deltas[p][j] + parent[j];
        }
        // checkAction: joint limits of every waypoint, then the edge parent -> waypoint
        if (!m_robot->checkJointLimits(succ)) {
            continue;
        }
        if (!m_cc->isStateToStateValid(parent, succ)) {
            continue;
        }
        stateToCoord(succ, coord);
        const int succ_id = getOrCreateState(coord, succ);
        const bool is_goal = isGoal(succ);
        succs.push_back(is_goal ? m_goal_state_id : succ_id);
        costs.push_back(m_prim_cost[p]);    // cost(parent, succ, weights[i], goal), manip_lattice.cpp:296
    }
}

/// mprimActive (manip_lattice_action_space.cpp:662-691)
bool ManipLatticePlanner::primActive(size_t p, bool near_goal) const
{
    if (!m_prim_short[p]) {
        return !(m_params.use_short_dist && near_goal);
    }
    return m_params.use_short_dist && near_goal;
}

/// ManipLattice::GetLazySuccs (manip_lattice.cpp:1012-1090): every action of the action space becomes a successor --
/// no limit test, no collision test -- at the optimistic cost, flagged "not the true cost"
void ManipLatticePlanner::getLazySuccs(int state_id, std::vector<int>& succs, std::vector<int>& costs,
                                       std::vector<bool>& true_costs)
{
    if (state_id == m_goal_state_id) {
        return; // goal state is absorbing
    }
    const std::vector<double> parent = m_states[state_id].state;
    std::vector<double> pose;
    m_robot->computePlanningLinkFK(parent, pose);
    const bool near_goal = m_heur->getMetricGoalDistance(pose[0], pose[1], pose[2]) <= m_params.short_dist_thresh;
    std::vector<double> succ(parent.size());
    std::vector<int> coord;
    for (size_t p = 0; p < m_prim_deltas.size(); ++p) {
        if (!primActive(p, near_goal)) {
            continue;
        }
        for (size_t j = 0; j < parent.size(); ++j) {
            succ[j] = m_prim_deltas[p][j] + parent[j];
        }
        stateToCoord(succ, coord);
        const bool is_goal = isGoal(succ);
        const int succ_id = getOrCreateState(coord, succ);
        succs.push_back(is_goal ? m_goal_state_id : succ_id);
        costs.push_back(1000);        // cost(): DefaultCostMultiplier (manip_lattice.cpp:1388-1412)
        true_costs.push_back(false);
    }
}

/// ManipLattice::GetTrueCost (manip_lattice.cpp:1094-1167): the cheapest valid action of the parent that ends in the
/// child's cell (or, for the goal id, in a goal state); -1 when there is none
int ManipLatticePlanner::getTrueCost(int parent_id, int child_id)
{
    const std::vector<double> parent = m_states[parent_id].state;
    std::vector<double> pose;
    m_robot->computePlanningLinkFK(parent, pose);
    const bool near_goal = m_heur->getMetricGoalDistance(pose[0], pose[1], pose[2]) <= m_params.short_dist_thresh;
    const bool goal_edge = child_id == m_goal_state_id;
    std::vector<double> succ(parent.size());
    std::vector<int> coord;
    int best_cost = std::numeric_limits<int>::max();
    for (size_t p = 0; p < m_prim_deltas.size(); ++p) {
        if (!primActive(p, near_goal)) {
            continue;
        }
        for (size_t j = 0; j < parent.size(); ++j) {
            succ[j] = m_prim_deltas[p][j] + parent[j];
        }
        stateToCoord(succ, coord);
        if (goal_edge) {
            if (!isGoal(succ)) {
                continue;
            }
        } else if (coord != m_states[child_id].coord) {
            continue;
        }
        if (!m_robot->checkJointLimits(succ) || !m_cc->isStateToStateValid(parent, succ)) {
            continue;   // checkAction
        }
        const int edge_cost = (int)(1000 * 1.0);   // cost(parent, succ, 1, goal_edge), :1414-1437
        if (edge_cost < best_cost) {
            best_cost = edge_cost;
        }
    }
    return best_cost != std::numeric_limits<int>::max() ? best_cost : -1;
}

/// setGoal + setStart of a query; false when the start is refused (res then holds the answer)
bool ManipLatticePlanner::begin(const std::vector<double>& start, const double goal_xyz[3], PlanResult& res)
{
    m_states.clear();
    m_coord_to_id.clear();
    for (int i = 0; i < 3; ++i) m_goal[i] = goal_xyz[i];
    m_goal_state_id = 0;                      // reserveHashEntry for the goal state (manip_lattice.cpp:122)
    m_states.push_back(LatticeState());
    m_heur->updateGoal(goal_xyz[0], goal_xyz[1], goal_xyz[2]);
    if (!m_robot->checkJointLimits(start) || !m_cc->isStateValid(start)) {   // setStart (manip_lattice.cpp:1944-1980)
        res.num_states = (int)m_states.size();
        return false;
    }
    std::vector<int> coord;
    stateToCoord(start, coord);
    m_start_state_id = getOrCreateState(coord, start);
    return true;
}

PlanResult ManipLatticePlanner::planLazy(const std::vector<double>& start, const double goal_xyz[3])
{
    PlanResult res;
    if (!begin(start, goal_xyz, res)) {
        return res;
    }
    const int bound = m_params.max_expansions;
    int expansions = 0, evaluations = 0;
    LazyAraStar search(
        [&](int id, std::vector<int>& succs, std::vector<int>& costs, std::vector<bool>& trues) {
            if (expansions >= bound) {
                return;
            }
            ++expansions;
            getLazySuccs(id, succs, costs, trues);
        },
        [&](int parent, int child) {
            if (expansions >= bound) {
                return -1;
            }
            ++evaluations;
            return getTrueCost(parent, child);
        },
        [this](int id) { return goalHeuristic(id); });
    const LazyAraStar::Result r = search.search(m_start_state_id, m_goal_state_id, m_params.epsilon);
    res.expansions = expansions;
    res.evaluations = evaluations;
    res.num_states = (int)m_states.size();
    if (!r.found) {
        return res;
    }
    res.path_ids = r.path;
    res.cost = r.cost;
    res.success = true;
    extractPath(res.path_ids, res.path_states);
    return res;
}

PlanResult ManipLatticePlanner::plan(const std::vector<double>& start, const double goal_xyz[3])
{
    PlanResult res;
    m_states.clear();
    m_coord_to_id.clear();
    for (int i = 0; i < 3; ++i) m_goal[i] = goal_xyz[i];

    // reserveHashEntry for the goal state (manip_lattice.cpp:122): id 0
    m_goal_state_id = 0;
    m_states.push_back(LatticeState());

    // setGoal -> BfsHeuristic::updateGoal
    m_heur->updateGoal(goal_xyz[0], goal_xyz[1], goal_xyz[2]);

    // setStart (manip_lattice.cpp:1944-1980)
    if (!m_robot->checkJointLimits(start) || !m_cc->isStateValid(start)) {
        res.num_states = (int)m_states.size();
        return res;
    }
    std::vector<int> coord;
    stateToCoord(start, coord);
    m_start_state_id = getOrCreateState(coord, start);

    // replan (arastar.cpp:107-215), first solution at the initial epsilon: oracle/arastar.h
    AraStar search(
        [this](int id, std::vector<int>& succs, std::vector<int>& costs) { getSuccs(id, succs, costs); },
        [this](int id) { return goalHeuristic(id); });
    const AraStar::Result r = search.search(m_start_state_id, m_goal_state_id, m_params.epsilon, m_params.max_expansions);
    res.expansions = r.expansions;
    res.num_states = (int)m_states.size();
    if (!r.found || r.cost >= AraStar::INFINITE_COST) {
        return res;   // no path, or the reference's degenerate "[goal] at INFINITECOST" answer (see arastar.h)
    }
    res.path_ids = r.path;
    res.cost = r.cost;
    res.success = true;
    extractPath(res.path_ids, res.path_states);
    return res;
}

/// ManipLattice::extractPath (manip_lattice.cpp:2018-2160): the joint values of every state of an id path.  The
/// goal id stands for any state satisfying the goal, so its predecessor's actions are applied again and the
/// cheapest valid one ending in a goal state (strict <, every action costs the same: the first) names the
/// lattice state whose joint values are reported.
bool ManipLatticePlanner::extractPath(const std::vector<int>& ids, std::vector<std::vector<double>>& path) const
{
    path.clear();
    if (ids.empty()) {
        return true;
    }
    if (ids.size() == 1) {
        path.push_back(m_states[ids[0] == m_goal_state_id ? m_start_state_id : ids[0]].state);
        return true;
    }
    if (ids[0] == m_goal_state_id) {
        return false;
    }
    path.push_back(m_states[ids[0]].state);
    for (size_t i = 1; i < ids.size(); ++i) {
        const int prev_id = ids[i - 1], curr_id = ids[i];
        if (prev_id == m_goal_state_id) {
            return false;
        }
        if (curr_id != m_goal_state_id) {
            path.push_back(m_states[curr_id].state);
            continue;
        }
        const std::vector<double>& parent = m_states[prev_id].state;
        std::vector<double> pose;
        m_robot->computePlanningLinkFK(parent, pose);
        const bool near_goal = m_heur->getMetricGoalDistance(pose[0], pose[1], pose[2]) <= m_params.short_dist_thresh;
        std::vector<double> succ(parent.size());
        std::vector<int> coord;
        int best_cost = std::numeric_limits<int>::max();
        int best = -1;
        for (size_t p = 0; p < m_prim_deltas.size(); ++p) {
            const bool active = m_prim_short[p] ? (m_params.use_short_dist && near_goal)
                                                : !(m_params.use_short_dist && near_goal);
            if (!active) {
                continue;
            }
            for (size_t j = 0; j < parent.size(); ++j) {
                succ[j] = m_prim_deltas[p][j] + parent[j];
            }
            if (!isGoal(succ)) {
                continue;
            }
            if (!m_robot->checkJointLimits(succ) || !m_cc->isStateToStateValid(parent, succ)) {
                continue;   // checkAction
            }
            stateToCoord(succ, coord);
            auto it = m_coord_to_id.find(coord);
            if (it == m_coord_to_id.end()) {
                continue;   // the reference asserts the entry exists
            }
            const int edge_cost = (int)(1000 * 1.0);
            if (edge_cost < best_cost) {
                best_cost = edge_cost;
                best = it->second;
            }
        }
        if (best < 0) {
            return false;
        }
        path.push_back(m_states[best].state);
    }
    return true;
}

} // namespace oracle
