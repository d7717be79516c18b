#include "batch_planner.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>

namespace smplhost {

static const int INFINITECOST = 1000000000;
static const int BFS_WALL = 0x7FFFFFFF;

namespace {
struct Timer
{
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double lap()
    {
        auto t1 = std::chrono::steady_clock::now();
        double s = std::chrono::duration<double>(t1 - t0).count();
        t0 = t1;
        return s;
    }
};

// smpl/angles.h:45-62
double normalizeAngle(double angle)
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
} // namespace

BatchPlanner::BatchPlanner(smplgpu_ctx* ctx, const PlannerConfig& cfg, int max_concurrent) :
    m_ctx(ctx), m_cfg(cfg), m_max_concurrent(std::max(1, max_concurrent))
{
    const int dof = cfg.dof;
    const int n_prims = (int)cfg.short_flags.size();
    // ManipLatticeActionSpace::addMotionPrim(..., add_converse = true)
    for (int p = 0; p < n_prims; ++p) {
        std::vector<double> d(cfg.mprims.begin() + (size_t)p * dof, cfg.mprims.begin() + (size_t)(p + 1) * dof);
        m_prim_deltas.push_back(d);
        m_prim_short.push_back(cfg.short_flags[p] != 0);
        for (double& v : d) v *= -1.0;
        m_prim_deltas.push_back(d);
        m_prim_short.push_back(cfg.short_flags[p] != 0);
    }
    // ManipLattice::init discretisation (manip_lattice.cpp:125-139)
    m_coord_vals.resize(dof);
    m_coord_deltas.resize(dof);
    for (int v = 0; v < dof; ++v) {
        const double res = cfg.resolutions[v];
        if (cfg.var_continuous[v]) {
            m_coord_vals[v] = (int)std::round((2.0 * M_PI) / res);
            m_coord_deltas[v] = (2.0 * M_PI) / (double)m_coord_vals[v];
        } else {
            const double span = std::fabs(cfg.var_max[v] - cfg.var_min[v]);
            m_coord_vals[v] = std::max(1, (int)std::round(span / res));
            m_coord_deltas[v] = span / (double)m_coord_vals[v];
        }
    }
}

// manip_lattice.cpp:1263-1289 (every KDL planning variable is either continuous or bounded)
void BatchPlanner::stateToCoord(const double* q, std::vector<int>& coord) const
{
    coord.resize(m_cfg.dof);
    for (int i = 0; i < m_cfg.dof; ++i) {
        if (m_cfg.var_continuous[i]) {
            double pos = normalizeAngle(q[i]);
            if (pos < 0.0) {
                pos += 2.0 * M_PI;
            }
            coord[i] = (int)((pos + m_coord_deltas[i] * 0.5) / m_coord_deltas[i]);
            if (coord[i] == m_coord_vals[i]) {
                coord[i] = 0;
            }
        } else {
            coord[i] = (int)(((q[i] - m_cfg.var_min[i]) / m_coord_deltas[i]) + 0.5);
        }
    }
}

// KDLRobotModel::checkJointLimits -> normalizeAnglesIntoRange (kdl_robot_model.cpp:173-235, 326-337)
bool BatchPlanner::checkJointLimits(const double* q) const
{
    for (int i = 0; i < m_cfg.dof; ++i) {
        if (m_cfg.var_min[i] > m_cfg.var_max[i]) {
            return false;
        }
    }
    for (int i = 0; i < m_cfg.dof; ++i) {
        const double a_min = m_cfg.var_min[i];
        const double a_max = normalizeAngle(m_cfg.var_min[i]); // sic: the reference passes normalize(min) as the upper wrap bound
        double a = q[i];
        if (std::fabs(a) > 2.0 * M_PI) {
            a = std::fmod(a, 2.0 * M_PI);
        }
        while (a > a_max) {
            a -= 2.0 * M_PI;
        }
        while (a < a_min) {
            a += 2.0 * M_PI;
        }
        if (a < m_cfg.var_min[i] || a > m_cfg.var_max[i]) {
            return false;
        }
    }
    return true;
}

// DistanceMap::worldToGrid (distance_map.hpp:520-527)
void BatchPlanner::worldToGrid(const double* p, int* cell) const
{
    const double inv = 1.0 / m_cfg.res;
    for (int a = 0; a < 3; ++a) {
        cell[a] = (int)(inv * (p[a] - (m_cfg.origin[a] - m_cfg.res)) + 0.5) - 1;
    }
}

int BatchPlanner::computeKey(const SState& s) const
{
    return s.g + (unsigned int)(m_cfg.epsilon * s.h);
}

BatchPlanner::SState& BatchPlanner::sstate(Query& Q, int id)
{
    if ((int)Q.search.size() <= id) {
        SState blank;
        blank.g = blank.h = blank.f = blank.eg = 0;
        blank.iteration_closed = 0;
        blank.bp = -1;
        blank.heap_index = 0;
        blank.touched = false;
        Q.search.resize(id + 1, blank);
    }
    return Q.search[id];
}

// ARAStar::reinitSearchState (arastar.cpp:613-627); one search call per query, so "reinit" == first touch
void BatchPlanner::touch(Query& Q, int id)
{
    SState& s = sstate(Q, id);
    if (!s.touched) {
        s.g = INFINITECOST;
        s.h = (id == 0) ? Q.goal_h : Q.states[id].h;
        s.f = INFINITECOST;
        s.eg = INFINITECOST;
        s.iteration_closed = 0;
        s.bp = -1;
        s.touched = true;
    }
}

void BatchPlanner::percolateUp(Query& Q, size_t pivot)
{
    const int tmp = Q.open[pivot];
    while (pivot != 1) {
        const size_t p = pivot >> 1;
        if (Q.search[Q.open[p]].f < Q.search[tmp].f) {
            break;
        }
        Q.open[pivot] = Q.open[p];
        Q.search[Q.open[pivot]].heap_index = (int)pivot;
        pivot = p;
    }
    Q.open[pivot] = tmp;
    Q.search[tmp].heap_index = (int)pivot;
}

void BatchPlanner::percolateDown(Query& Q, size_t pivot)
{
    if (pivot >= Q.open.size()) {
        return;
    }
    size_t left = pivot << 1, right = left + 1;
    const int tmp = Q.open[pivot];
    while (left < Q.open.size()) {
        size_t s = right;
        if (right >= Q.open.size() || Q.search[Q.open[left]].f < Q.search[Q.open[right]].f) {
            s = left;
        }
        if (Q.search[Q.open[s]].f < Q.search[tmp].f) {
            Q.open[pivot] = Q.open[s];
            Q.search[Q.open[pivot]].heap_index = (int)pivot;
            pivot = s;
        } else {
            break;
        }
        left = pivot << 1;
        right = left + 1;
    }
    Q.open[pivot] = tmp;
    Q.search[tmp].heap_index = (int)pivot;
}

void BatchPlanner::heapPush(Query& Q, int id)
{
    Q.search[id].heap_index = (int)Q.open.size();
    Q.open.push_back(id);
    percolateUp(Q, Q.open.size() - 1);
}

void BatchPlanner::heapPop(Query& Q)
{
    Q.search[Q.open[1]].heap_index = 0;
    Q.open[1] = Q.open.back();
    Q.open.pop_back();
    percolateDown(Q, 1);
}

void BatchPlanner::finish(Query& Q, bool found)
{
    Q.done = true;
    Q.result.num_states = (int)Q.states.size();
    if (!found || Q.search.empty() || Q.search[0].g >= INFINITECOST) {
        return;
    }
    for (int s = 0; s >= 0; s = Q.search[s].bp) {
        Q.result.path_ids.push_back(s);
    }
    std::reverse(Q.result.path_ids.begin(), Q.result.path_ids.end());
    Q.result.cost = Q.search[0].g;
    Q.result.success = true;
}

bool BatchPlanner::plan(const double* starts, const double* goals, int nq, std::vector<QueryResult>& out, std::string* err)
{
    out.assign(nq, QueryResult());
    m_stats = BatchStats();
    Timer t;
    const int slots = std::min(m_max_concurrent, std::max(1, nq));
    int r = smplgpu_bfs_bank_create(m_ctx, slots, m_cfg.inflation_radius);
    m_stats.device_seconds += t.lap();
    ++m_stats.device_calls;
    if (r < 0) {
        if (err) *err = smplgpu_last_error(m_ctx);
        return false;
    }
    for (int first = 0; first < nq; first += slots) {
        std::vector<int> ids;
        for (int i = first; i < std::min(nq, first + slots); ++i) ids.push_back(i);
        if (!runWave(starts, goals, ids, out, err)) {
            return false;
        }
    }
    return true;
}

bool BatchPlanner::runWave(const double* starts, const double* goals, const std::vector<int>& ids,
                           std::vector<QueryResult>& out, std::string* err)
{
    const int dof = m_cfg.dof;
    const int nw = (int)ids.size();
    Timer t;
    auto fail_dev = [&]() {
        if (err) *err = smplgpu_last_error(m_ctx);
        return false;
    };

    // ---- setGoal: one BFS per query, all in one bank run (BfsHeuristic::updateGoal) ----
    std::vector<Query> W(nw);
    const int bank_slots = std::min(m_max_concurrent, std::max(1, (int)out.size()));
    std::vector<int32_t> seeds((size_t)bank_slots * 3, -1);
    for (int k = 0; k < nw; ++k) {
        Query& Q = W[k];
        Q.index = ids[k];
        Q.slot = k;
        Q.done = false;
        Q.expanding = -1;
        for (int a = 0; a < 3; ++a) Q.goal[a] = goals[(size_t)ids[k] * 3 + a];
        int cell[3];
        worldToGrid(Q.goal, cell);
        const bool inb = cell[0] >= 0 && cell[1] >= 0 && cell[2] >= 0 &&
                         cell[0] < m_cfg.dims[0] && cell[1] < m_cfg.dims[1] && cell[2] < m_cfg.dims[2];
        for (int a = 0; a < 3; ++a) seeds[(size_t)k * 3 + a] = cell[a];
        // the seed cell holds distance 0 even if it was a wall (bfs3d.cpp:181-187)
        Q.goal_h = inb ? 0 : SMPLGPU_HEURISTIC_INFINITY;
        Q.open.assign(1, -1);
        Q.states.push_back(LState()); // id 0 = the goal state (manip_lattice.cpp:122)
    }
    m_stats.host_seconds += t.lap();
    if (smplgpu_bfs_bank_run(m_ctx, seeds.data()) < 0) return fail_dev();
    ++m_stats.device_calls;

    // ---- setStart: limits + validity, then heuristic / metric distance of the start ----
    std::vector<double> q0((size_t)nw * dof), q1;
    std::vector<int32_t> slot(nw);
    for (int k = 0; k < nw; ++k) {
        std::copy(starts + (size_t)ids[k] * dof, starts + (size_t)(ids[k] + 1) * dof, q0.begin() + (size_t)k * dof);
        slot[k] = k;
    }
    std::vector<uint8_t> verdict(nw);
    std::vector<int32_t> h(nw), gd(nw);
    std::vector<double> off((size_t)nw * 3);
    if (smplgpu_is_states_valid(m_ctx, q0.data(), nw, verdict.data()) < 0) return fail_dev();
    std::vector<uint8_t> dummy(nw);
    if (smplgpu_expand_batch(m_ctx, q0.data(), q0.data(), slot.data(), nw, m_cfg.cost_per_cell,
                             dummy.data(), h.data(), gd.data(), off.data()) < 0) return fail_dev();
    m_stats.device_calls += 2;
    m_stats.device_seconds += t.lap();

    int active = 0;
    std::vector<int> coord;
    for (int k = 0; k < nw; ++k) {
        Query& Q = W[k];
        const double* qs = &q0[(size_t)k * dof];
        if (!checkJointLimits(qs) || !verdict[k]) {
            finish(Q, false);
            continue;
        }
        stateToCoord(qs, coord);
        LState ls;
        ls.coord = coord;
        ls.q.assign(qs, qs + dof);
        ls.h = h[k];
        ls.gdist = gd[k];
        Q.states.push_back(ls);
        Q.coord_to_id[coord] = 1;
        sstate(Q, 1);
        touch(Q, 1);
        touch(Q, 0);
        Q.search[1].g = 0;
        Q.search[1].f = computeKey(Q.search[1]);
        heapPush(Q, 1);
        ++active;
    }

    // ---- lock-step rounds ----
    struct EdgeRef { int k; };
    std::vector<EdgeRef> owner;
    while (active > 0) {
        q0.clear();
        q1.clear();
        slot.clear();
        owner.clear();
        for (int k = 0; k < nw; ++k) {
            Query& Q = W[k];
            if (Q.done) {
                continue;
            }
            Q.expanding = -1;
            if (Q.open.size() <= 1) {
                finish(Q, false);
                --active;
                continue;
            }
            const int min_id = Q.open[1];
            if (Q.search[min_id].f >= Q.search[0].f || min_id == 0) {
                finish(Q, true);
                --active;
                continue;
            }
            if (Q.result.expansions >= m_cfg.max_expansions) {
                finish(Q, false);
                --active;
                continue;
            }
            heapPop(Q);
            Q.search[min_id].iteration_closed = 1;
            Q.search[min_id].eg = Q.search[min_id].g;
            Q.expanding = min_id;
            ++Q.result.expansions;

            // ManipLatticeActionSpace::apply: which primitives are active at this state
            const LState& P = Q.states[min_id];
            const double goal_dist = (double)P.gdist * m_cfg.res;
            const bool near_goal = goal_dist <= m_cfg.short_dist_thresh;
            for (size_t p = 0; p < m_prim_deltas.size(); ++p) {
                const bool active_prim = m_prim_short[p] ? (m_cfg.use_short_dist && near_goal)
                                                         : !(m_cfg.use_short_dist && near_goal);
                if (!active_prim) {
                    continue;
                }
                const size_t base = q1.size();
                q1.resize(base + dof);
                for (int j = 0; j < dof; ++j) {
                    q1[base + j] = m_prim_deltas[p][j] + P.q[j];
                }
                if (!checkJointLimits(&q1[base])) { // checkAction: joint limits first
                    q1.resize(base);
                    continue;
                }
                q0.insert(q0.end(), P.q.begin(), P.q.end());
                slot.push_back(Q.slot);
                owner.push_back(EdgeRef{ k });
            }
        }
        const int ne = (int)owner.size();
        m_stats.host_seconds += t.lap();
        if (ne == 0) {
            continue; // queries that expanded a state without any in-limits successor carry on next round
        }
        verdict.resize(ne);
        h.resize(ne);
        gd.resize(ne);
        off.resize((size_t)ne * 3);
        if (smplgpu_expand_batch(m_ctx, q0.data(), q1.data(), slot.data(), ne, m_cfg.cost_per_cell,
                                 verdict.data(), h.data(), gd.data(), off.data()) < 0) return fail_dev();
        ++m_stats.device_calls;
        ++m_stats.rounds;
        m_stats.edges_submitted += ne;
        m_stats.device_seconds += t.lap();

        // ---- GetSuccs bookkeeping + ARAStar::expand relaxations, per query in submission order ----
        for (int e = 0; e < ne; ++e) {
            if (!verdict[e]) {
                continue;
            }
            Query& Q = W[owner[e].k];
            const double* qs = &q1[(size_t)e * dof];
            stateToCoord(qs, coord);
            int succ_id;
            auto it = Q.coord_to_id.find(coord);
            if (it != Q.coord_to_id.end()) {
                succ_id = it->second;
            } else {
                succ_id = (int)Q.states.size();
                LState ls;
                ls.coord = coord;
                ls.q.assign(qs, qs + dof);
                ls.h = h[e];
                ls.gdist = gd[e];
                Q.states.push_back(ls);
                Q.coord_to_id[coord] = succ_id;
            }
            // ManipLattice::isGoal, XYZ_GOAL (manip_lattice.cpp:1673-1687)
            const bool is_goal = std::fabs(off[3 * e] - Q.goal[0]) <= m_cfg.xyz_tolerance[0] &&
                                 std::fabs(off[3 * e + 1] - Q.goal[1]) <= m_cfg.xyz_tolerance[1] &&
                                 std::fabs(off[3 * e + 2] - Q.goal[2]) <= m_cfg.xyz_tolerance[2];
            const int target = is_goal ? 0 : succ_id;
            sstate(Q, target);
            touch(Q, target);
            SState& ss = Q.search[target];
            const int new_cost = Q.search[Q.expanding].eg + (int)(1000 * 1.0);
            if (new_cost < ss.g) {
                ss.g = new_cost;
                ss.bp = Q.expanding;
                if (ss.iteration_closed != 1) {
                    ss.f = computeKey(ss);
                    if (ss.heap_index != 0) {
                        percolateUp(Q, (size_t)ss.heap_index);
                    } else {
                        heapPush(Q, target);
                    }
                }
            }
        }
        m_stats.host_seconds += t.lap();
    }
    for (int k = 0; k < nw; ++k) {
        out[W[k].index] = W[k].result;
    }
    return true;
}

} // namespace smplhost
