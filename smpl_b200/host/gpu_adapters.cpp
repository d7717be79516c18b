// Adapters of gpu_adapters.h: each virtual of the reference's plugin interfaces becomes one call of the
// C ABI (n = 1), the batched entry points submit everything in one call.  No arithmetic of the hot path
// lives here: verdicts, heuristics and FK come from the device.
#include "gpu_adapters.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace smplhost {

using sbpl::motion::Extension;
using sbpl::motion::GetClassCode;
using sbpl::motion::RobotState;

///////////////////////////////////////////////////////////////////////////////
// ExpansionCache
///////////////////////////////////////////////////////////////////////////////

ExpansionCache::ExpansionCache(smplgpu_ctx* ctx, int dof, const double* deltas, int n_prims, int cost_per_cell) :
    m_ctx(ctx), m_dof(dof), m_prims(n_prims), m_cost_per_cell(cost_per_cell),
    m_deltas(deltas, deltas + (size_t)std::max(0, n_prims) * (size_t)std::max(0, dof))
{
    m_ok = ctx != nullptr && dof > 0 && dof <= SMPLGPU_MAX_DOF && n_prims >= 0 &&
           smplgpu_set_motion_primitives(ctx, m_deltas.data(), n_prims) == 0;
}

bool ExpansionCache::valid()
{
    if (!m_ok) {
        return false;
    }
    if (m_n > 0 && smplgpu_scene_epoch(m_ctx) != m_epoch) {
        m_n = 0;   // the scene, the walls or the goal changed: earlier answers may no longer hold
    }
    return true;
}

bool ExpansionCache::expand(const double* q)
{
    m_n = 0;
    const smplgpu_succ_info* info = nullptr;
    if (smplgpu_expand_state(m_ctx, q, m_cost_per_cell, &info) != 0 || info == nullptr) {
        return false;
    }
    ++m_launches;
    m_info = info;
    m_n = m_prims + 1;
    m_hint = 0;
    m_has_pending = false;
    m_epoch = smplgpu_scene_epoch(m_ctx);
    return true;
}

// b == a + delta for some primitive (the addition the device does)
bool ExpansionCache::isSuccessorOf(const double* a, const double* b) const
{
    for (int p = 0; p < m_prims; ++p) {
        const double* d = &m_deltas[(size_t)p * m_dof];
        bool same = true;
        for (int j = 0; j < m_dof && same; ++j) {
            const double s = d[j] + a[j];
            same = std::memcmp(&s, &b[j], sizeof(double)) == 0;
        }
        if (same) {
            return true;
        }
    }
    return false;
}

const smplgpu_succ_info* ExpansionCache::find(const double* q)
{
    if (!valid() || m_n == 0) {
        return nullptr;
    }
    const size_t bytes = (size_t)m_dof * sizeof(double);
    // the caller walks the successors in table order: try where the last hit was first, then onwards
    for (int k = 0; k < m_n; ++k) {
        const int i = (m_hint + k) % m_n;
        if (std::memcmp(m_info[i].state, q, bytes) == 0) {
            m_hint = i;
            ++m_hits;
            if (i > 0) {
                m_pending.assign(q, q + m_dof);
                m_has_pending = true;
            }
            return &m_info[i];
        }
    }
    return nullptr;
}

const smplgpu_succ_info* ExpansionCache::get(const double* q)
{
    if (const smplgpu_succ_info* r = find(q)) {
        return r;
    }
    if (!m_ok) {
        return nullptr;
    }
    // A search that expands a state right after its parent asks about that state first -- answered from the parent's
    // record, as a successor entry -- and then about ITS successors, which no record holds yet: the state to expand is
    // the one found last, not the successor asked about (a lazy search asks nothing else in between: without this,
    // GetLazySuccs cost one launch per successor).
    if (m_has_pending && isSuccessorOf(m_pending.data(), q)) {
        const std::vector<double> parent_state = m_pending;
        if (!expand(parent_state.data())) {
            return nullptr;
        }
        if (const smplgpu_succ_info* r = find(q)) {
            return r;
        }
    }
    if (!expand(q)) {
        return nullptr;
    }
    return &m_info[0];
}

const smplgpu_succ_info* ExpansionCache::parent(const double* q)
{
    if (!valid()) {
        return nullptr;
    }
    if (m_n > 0 && std::memcmp(m_info[0].state, q, (size_t)m_dof * sizeof(double)) == 0) {
        ++m_hits;
        return &m_info[0];
    }
    return expand(q) ? &m_info[0] : nullptr;
}

const smplgpu_succ_info* ExpansionCache::edge(const double* a, const double* b)
{
    if (!valid()) {
        return nullptr;
    }
    const size_t bytes = (size_t)m_dof * sizeof(double);
    if (m_n == 0 || std::memcmp(m_info[0].state, a, bytes) != 0) {
        // a is not the current parent: expanding it pays off only if b is one of its successors, i.e.
        // b == a + delta for some primitive (the addition the device will do)
        const bool is_succ = isSuccessorOf(a, b);
        if (!is_succ || !expand(a)) {
            return nullptr;
        }
    }
    for (int k = 0; k < m_prims; ++k) {
        const int i = 1 + (std::max(0, m_hint - 1) + k) % m_prims;
        if (std::memcmp(m_info[i].state, b, bytes) == 0) {
            m_hint = i;
            ++m_hits;
            return &m_info[i];
        }
    }
    return nullptr;
}

const smplgpu_succ_info* ExpansionCache::findLink(double x, double y, double z)
{
    if (!valid() || m_n == 0) {
        return nullptr;
    }
    const double p[3] = { x, y, z };
    for (int k = 0; k < m_n; ++k) {
        const int i = (m_hint + k) % m_n;
        if (std::memcmp(m_info[i].link_xyz, p, sizeof(p)) == 0) {
            ++m_hits;
            return &m_info[i];
        }
    }
    return nullptr;
}

///////////////////////////////////////////////////////////////////////////////
// GpuCollisionSpace
///////////////////////////////////////////////////////////////////////////////

// CollisionSpace::isStateValid (collision_space.cpp:532-536)
bool GpuCollisionSpace::isStateValid(const RobotState& state, bool)
{
    if ((int)state.size() != m_dof) {
        return false;
    }
    if (m_cache) {
        if (const smplgpu_succ_info* r = m_cache->parent(state.data())) {
            return r->state_valid != 0;
        }
    }
    uint8_t v = 0;
    if (smplgpu_is_states_valid(m_ctx, state.data(), 1, &v) != 0) {
        return false; // errors are `false`, never exceptions (collision_checker.h convention)
    }
    return v != 0;
}

// collision_space.h:202-205 has an empty body (SURVEY.md section 8a defect 3): return the plain overload's
// verdict and dist = +max
bool GpuCollisionSpace::isStateValid(const RobotState& state, double& distToObst, bool verbose)
{
    distToObst = std::numeric_limits<double>::max();
    return isStateValid(state, verbose);
}

// CollisionSpace::isStateToStateValid (collision_space.cpp:538-581)
bool GpuCollisionSpace::isStateToStateValid(const RobotState& start, const RobotState& finish, bool)
{
    if ((int)start.size() != m_dof || (int)finish.size() != m_dof) {
        return false;
    }
    if (m_cache) {
        if (const smplgpu_succ_info* r = m_cache->edge(start.data(), finish.data())) {
            return r->edge_valid != 0;
        }
    }
    uint8_t v = 0;
    if (smplgpu_is_edges_valid(m_ctx, start.data(), finish.data(), 1, &v, nullptr) != 0) {
        return false;
    }
    return v != 0;
}

bool GpuCollisionSpace::isStateToStateValid(const RobotState& a0, const RobotState& a1, double& distToObst,
                                            int& distToObstCells, bool verbose)
{
    distToObst = std::numeric_limits<double>::max();
    distToObstCells = std::numeric_limits<int>::max();
    return isStateToStateValid(a0, a1, verbose);
}

static double normalizeAngle(double angle)
{
    if (std::fabs(angle) > 2.0 * M_PI) {
        angle = std::fmod(angle, 2.0 * M_PI);
    }
    if (angle < -M_PI) {
        angle += 2.0 * M_PI;
    }
    if (angle > M_PI) {
        angle -= 2.0 * M_PI;
    }
    return angle;
}

// CollisionSpace::interpolatePath (collision_space.cpp:583-612): the waypoints isStateToStateValid checks
// (robot_motion_collision_model.h:173-181, 224-249, 297-321; .cpp:371-407).  Post-processing only, so it
// stays on the host.  The reference's limit test is inverted (it rejects every motion whose end points are
// WITHIN the limits, collision_space.cpp:591-596, SURVEY.md section 8a defect 7, so its InterpolatePath fails on
// every legal path); NO limit test is made here -- a deliberate deviation, recorded in DESIGN.md -- callers that
// need one ask GpuRobotModel::checkJointLimits.
bool GpuCollisionSpace::interpolatePath(const RobotState& start, const RobotState& finish,
                                        std::vector<RobotState>& path)
{
    if ((int)start.size() != m_dof || (int)finish.size() != m_dof || (int)m_types.size() != m_dof) {
        return false;
    }
    double motion = 0.0;
    std::vector<double> diffs(m_dof);
    for (int v = 0; v < m_dof; ++v) {
        if (m_types[v] == SMPLGPU_VAR_CONTINUOUS) {
            diffs[v] = normalizeAngle(finish[v] - start[v]);
            motion += m_weights[v] * std::fabs(diffs[v]);
        } else if (m_types[v] == SMPLGPU_VAR_REVOLUTE) {
            diffs[v] = finish[v] - start[v];
            motion += m_weights[v] * std::fabs(diffs[v]);
        } else {
            diffs[v] = finish[v] - start[v];
            motion += std::fabs(diffs[v]);
        }
    }
    int count = 0;
    if (motion != 0.0) {
        count = std::max(2, (int)std::ceil(motion / 0.05) + 1);
    }
    path.resize(count);
    const double inv = count > 1 ? 1.0 / (double)(count - 1) : 0.0;
    for (int i = 0; i < count; ++i) {
        const double alpha = (double)i * inv;
        path[i].resize(m_dof);
        for (int v = 0; v < m_dof; ++v) {
            path[i][v] = start[v] + alpha * diffs[v];
        }
    }
    return true;
}

Extension* GpuCollisionSpace::getExtension(size_t class_code)
{
    if (class_code == GetClassCode<sbpl::motion::CollisionChecker>()) {
        return static_cast<sbpl::motion::CollisionChecker*>(this);
    }
    if (class_code == GetClassCode<sbpl::motion::CollisionDistanceExtension>()) {
        return static_cast<sbpl::motion::CollisionDistanceExtension*>(this);
    }
    return nullptr;
}

// CollisionSpace::collisionDistance (collision_space.cpp:496-500); errors give 0 (no clearance), never exceptions
double GpuCollisionSpace::distanceToCollision(const RobotState& state)
{
    if ((int)state.size() != m_dof) {
        return 0.0;
    }
    double d = 0.0;
    if (smplgpu_collision_distance(m_ctx, state.data(), 1, &d) != 0) {
        return 0.0;
    }
    return d;
}

double GpuCollisionSpace::distanceToCollision(const RobotState& start, const RobotState& finish)
{
    std::vector<RobotState> path;
    if (!interpolatePath(start, finish, path)) {
        return 0.0;
    }
    if (path.empty()) {
        return distanceToCollision(start);   // zero-length motion
    }
    std::vector<double> d;
    if (!collisionDistances(path, d)) {
        return 0.0;
    }
    double best = d[0];
    for (double v : d) best = std::min(best, v);
    return best;
}

static bool flatten(const std::vector<RobotState>& states, int dof, std::vector<double>& buf)
{
    buf.resize(states.size() * (size_t)dof);
    for (size_t i = 0; i < states.size(); ++i) {
        if ((int)states[i].size() != dof) {
            return false;
        }
        std::copy(states[i].begin(), states[i].end(), buf.begin() + i * dof);
    }
    return true;
}

bool GpuCollisionSpace::isStatesValid(const std::vector<RobotState>& states, std::vector<uint8_t>& valid)
{
    valid.assign(states.size(), 0);
    if (states.empty()) {
        return true;
    }
    if (!flatten(states, m_dof, m_buf0)) {
        return false;
    }
    return smplgpu_is_states_valid(m_ctx, m_buf0.data(), (int)states.size(), valid.data()) == 0;
}

bool GpuCollisionSpace::collisionDistances(const std::vector<RobotState>& states, std::vector<double>& dist)
{
    dist.assign(states.size(), 0.0);
    if (states.empty()) {
        return true;
    }
    if (!flatten(states, m_dof, m_buf0)) {
        return false;
    }
    return smplgpu_collision_distance(m_ctx, m_buf0.data(), (int)states.size(), dist.data()) == 0;
}

bool GpuCollisionSpace::isEdgesValid(const std::vector<RobotState>& starts, const std::vector<RobotState>& finishes,
                                     std::vector<uint8_t>& valid)
{
    valid.assign(starts.size(), 0);
    if (starts.size() != finishes.size()) {
        return false;
    }
    if (starts.empty()) {
        return true;
    }
    if (!flatten(starts, m_dof, m_buf0) || !flatten(finishes, m_dof, m_buf1)) {
        return false;
    }
    return smplgpu_is_edges_valid(m_ctx, m_buf0.data(), m_buf1.data(), (int)starts.size(), valid.data(), nullptr) == 0;
}

///////////////////////////////////////////////////////////////////////////////
// GpuRobotModel
///////////////////////////////////////////////////////////////////////////////

GpuRobotModel::GpuRobotModel(smplgpu_ctx* ctx, RobotTables* tables, const std::string& planning_link) :
    m_ctx(ctx), m_planning_link(planning_link),
    m_min(tables->varMin()), m_max(tables->varMax()), m_cont(tables->varContinuous())
{
}

// KDLRobotModel::checkJointLimits (kdl_robot_model.cpp:326-337)
bool GpuRobotModel::checkJointLimits(const RobotState& state, bool)
{
    if (state.size() != m_min.size()) {
        return false;
    }
    if (m_cache) {
        if (const smplgpu_succ_info* r = m_cache->get(state.data())) {
            return r->limits_ok != 0;
        }
    }
    uint8_t ok = 0;
    if (smplgpu_check_joint_limits(m_ctx, state.data(), 1, &ok) != 0) {
        return false;
    }
    return ok != 0;
}

// KDLRobotModel::computeFK (kdl_robot_model.cpp:362-398): only the planning link's frame is on the hot path
bool GpuRobotModel::computeFK(const RobotState& state, const std::string& name, std::vector<double>& pose)
{
    if (name != m_planning_link) {
        return false;
    }
    return computePlanningLinkFK(state, pose);
}

// KDLRobotModel::computePlanningLinkFK (kdl_robot_model.cpp:400-423) + getTargetOffsetPose
bool GpuRobotModel::computePlanningLinkFK(const RobotState& state, std::vector<double>& pose)
{
    if (state.size() != m_min.size()) {
        return false;
    }
    pose.resize(6);
    if (m_cache) {
        if (const smplgpu_succ_info* r = m_cache->get(state.data())) {
            std::copy(r->pose, r->pose + 6, pose.begin());
            return true;
        }
    }
    return smplgpu_planning_frame_fk(m_ctx, state.data(), 1, pose.data()) == 0;
}

Extension* GpuRobotModel::getExtension(size_t class_code)
{
    if (class_code == GetClassCode<sbpl::motion::RobotModel>() ||
        class_code == GetClassCode<sbpl::motion::ForwardKinematicsInterface>()) {
        return this;
    }
    return nullptr;
}

///////////////////////////////////////////////////////////////////////////////
// GpuBfsHeuristic
///////////////////////////////////////////////////////////////////////////////

GpuBfsHeuristic::GpuBfsHeuristic(smplgpu_ctx* ctx, const double origin[3], double res, const int dims[3]) :
    m_ctx(ctx), m_res(res)
{
    for (int a = 0; a < 3; ++a) {
        m_origin[a] = origin[a];
        m_dims[a] = dims[a];
    }
}

// BfsHeuristic::init -> syncGridAndBfs (bfs_heuristic.cpp:45-71, 331-353): walls from the distance field
bool GpuBfsHeuristic::init(StateLookup lookup, int goal_state_id)
{
    m_lookup = lookup;
    m_goal_state_id = goal_state_id;
    const int walls = smplgpu_bfs_set_walls_from_df(m_ctx, m_inflation_radius);
    if (walls < 0) {
        return false;
    }
    m_walls = walls;
    return true;
}

// DistanceMap::worldToGrid (distance_map.hpp:520-527)
void GpuBfsHeuristic::worldToGrid(double x, double y, double z, int cell[3]) const
{
    const double inv = 1.0 / m_res;
    const double p[3] = { x, y, z };
    for (int a = 0; a < 3; ++a) {
        cell[a] = (int)(inv * (p[a] - (m_origin[a] - m_res)) + 0.5) - 1;
    }
}

// bfs_heuristic.cpp:83-101
void GpuBfsHeuristic::updateGoal(const sbpl::motion::GoalConstraint& goal)
{
    if (goal.tgt_off_pose.size() < 3) {
        return;
    }
    for (int a = 0; a < 3; ++a) m_goal_xyz[a] = goal.tgt_off_pose[a];
    int32_t cell[3];
    worldToGrid(m_goal_xyz[0], m_goal_xyz[1], m_goal_xyz[2], cell);
    smplgpu_bfs_run(m_ctx, cell, 1); // an out-of-bounds goal leaves every free cell undiscovered, as BFS_3D::run does
}

// BFS_3D::getDistance through the ABI; -2 = out of bounds
int GpuBfsHeuristic::cellCost(const int cell[3])
{
    int32_t c[3] = { cell[0], cell[1], cell[2] };
    int32_t d = -2;
    if (smplgpu_bfs_distances(m_ctx, c, 1, &d) != 0) {
        return -2;
    }
    return d;
}

// bfs_heuristic.cpp:103-125 needs the start state's projection, which only the planning space knows.
// ManipLatticeActionSpace::apply does call it on every expansion (manip_lattice_action_space.cpp:396), but this
// fork never uses the result (its only reader, the near_start test, is commented out), so 0 changes nothing.
double GpuBfsHeuristic::getMetricStartDistance(double, double, double)
{
    return 0.0;
}

// bfs_heuristic.cpp:127-138
double GpuBfsHeuristic::getMetricGoalDistance(double x, double y, double z)
{
    int d;
    const smplgpu_succ_info* r = m_cache ? m_cache->findLink(x, y, z) : nullptr;
    if (r != nullptr) {
        d = r->goal_dist_cells;   // the BFS value at the cell of the planning link of a state in the record
    } else {
        int cell[3];
        worldToGrid(x, y, z, cell);
        d = cellCost(cell);
    }
    if (d == -2) {
        return (double)0x7FFFFFFF * m_res;
    }
    return (double)d * m_res;
}

// bfs_heuristic.cpp:148-163: 0 when the state cannot be projected; the goal state projects to the goal pose
int GpuBfsHeuristic::GetGoalHeuristic(int state_id)
{
    if (state_id == m_goal_state_id) {
        int cell[3];
        worldToGrid(m_goal_xyz[0], m_goal_xyz[1], m_goal_xyz[2], cell);
        const int d = cellCost(cell);
        if (d == -2 || d == 0x7FFFFFFF) {
            return Infinity;
        }
        return m_cost_per_cell * d;
    }
    RobotState q;
    if (!m_lookup || !m_lookup(state_id, q)) {
        return 0;
    }
    if (m_cache && (int)q.size() <= SMPLGPU_MAX_DOF) {
        if (const smplgpu_succ_info* r = m_cache->get(q.data())) {
            return r->h;
        }
    }
    int32_t h = 0;
    if (smplgpu_goal_heuristics(m_ctx, q.data(), 1, m_cost_per_cell, &h) != 0) {
        return 0;
    }
    return h;
}

bool GpuBfsHeuristic::goalHeuristics(const std::vector<RobotState>& states, std::vector<int>& h)
{
    h.assign(states.size(), 0);
    if (states.empty()) {
        return true;
    }
    std::vector<double> buf;
    if (!flatten(states, (int)states[0].size(), buf)) {
        return false;
    }
    std::vector<int32_t> out(states.size());
    if (smplgpu_goal_heuristics(m_ctx, buf.data(), (int)states.size(), m_cost_per_cell, out.data()) != 0) {
        return false;
    }
    h.assign(out.begin(), out.end());
    return true;
}

Extension* GpuBfsHeuristic::getExtension(size_t class_code)
{
    if (class_code == GetClassCode<sbpl::motion::RobotHeuristic>()) {
        return this;
    }
    return nullptr;
}

} // namespace smplhost
