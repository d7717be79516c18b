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

int BatchPlanner::Lattice::find(const int* c) const
{
    if (table.empty()) {
        return -1;
    }
    const size_t mask = table.size() - 1;
    for (size_t i = (size_t)hash(c, dof) & mask;; i = (i + 1) & mask) {
        const int id = table[i];
        if (id < 0) {
            return -1;
        }
        if (std::equal(c, c + dof, &coords[(size_t)id * dof])) {
            return id;
        }
    }
}

void BatchPlanner::Lattice::grow()
{
    const size_t n = table.empty() ? 256 : table.size() * 2;
    table.assign(n, -1);
    const size_t mask = n - 1;
    for (int id = 1; id < size(); ++id) {   // id 0 (the goal state) has no coordinate
        size_t i = (size_t)hash(&coords[(size_t)id * dof], dof) & mask;
        while (table[i] >= 0) {
            i = (i + 1) & mask;
        }
        table[i] = id;
    }
}

int BatchPlanner::Lattice::add(const int* c, const double* q, int hv, int gd, bool index)
{
    const int id = size();
    if (c != nullptr) {
        coords.insert(coords.end(), c, c + dof);
        qs.insert(qs.end(), q, q + dof);
    } else {
        coords.insert(coords.end(), dof, 0);
        qs.insert(qs.end(), dof, 0.0);
    }
    h.push_back(hv);
    gdist.push_back(gd);
    if (index) {
        if ((size_t)(id + 1) * 2 > table.size()) {
            grow();   // re-enters every state including this one
        } else {
            const size_t mask = table.size() - 1;
            size_t i = (size_t)hash(c, dof) & mask;
            while (table[i] >= 0) {
                i = (i + 1) & mask;
            }
            table[i] = id;
        }
    }
    return id;
}

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
        s.h = (id == 0) ? Q.goal_h : Q.lat.h[id];
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
    Q.result.num_states = Q.lat.size();
    if (!found || Q.search.empty() || Q.search[0].g >= (unsigned int)INFINITECOST) {
        return;
    }
    for (int s = 0; s >= 0; s = Q.search[s].bp) {
        Q.result.path_ids.push_back(s);
    }
    std::reverse(Q.result.path_ids.begin(), Q.result.path_ids.end());
    Q.result.cost = Q.search[0].g;
    Q.result.success = true;
    // ManipLattice::extractPath (manip_lattice.cpp:2018-2160)
    const int dof = m_cfg.dof;
    const std::vector<int>& ids = Q.result.path_ids;
    Q.result.path_states.assign(ids.size() * (size_t)dof, 0.0);
    for (size_t i = 0; i < ids.size(); ++i) {
        int id = ids[i];
        if (id == 0) {   // the goal state id stands for "any state satisfying the goal"
            id = -1;
            if (i > 0) {
                for (const std::pair<int, int>& gs : Q.goal_succ) {
                    if (gs.first == ids[i - 1]) {
                        id = gs.second;
                        break;
                    }
                }
            }
            if (id < 0) {
                Q.result.path_states.resize(i * (size_t)dof);   // cannot happen for a found path; keep what is known
                break;
            }
        }
        std::copy(Q.lat.q(id), Q.lat.q(id) + dof, Q.result.path_states.begin() + i * (size_t)dof);
    }
}

///////////////////////////////////////////////////////////////////////////////
// fork-join pool
///////////////////////////////////////////////////////////////////////////////

BatchPlanner::Pool::Pool(int n) : m_n(std::max(1, n))
{
    for (int t = 1; t < m_n; ++t) {
        m_threads.emplace_back(&Pool::worker, this, t);
    }
}

BatchPlanner::Pool::~Pool()
{
    {
        std::lock_guard<std::mutex> lk(m_mutex);
        m_stop.store(true);
        m_generation.fetch_add(1);
    }
    m_cv.notify_all();
    for (std::thread& t : m_threads) {
        t.join();
    }
}

// workers spin briefly for the next job (jobs arrive every few tens of microseconds while the planner is
// busy) and block on the condition variable when nothing comes, so idle workers do not eat the cores the
// other ranks' planners need
void BatchPlanner::Pool::worker(int tid)
{
    int seen = 0;
    for (;;) {
        int spins = 0;
        while (m_generation.load(std::memory_order_acquire) == seen) {
            if (++spins < 4000) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            } else {
                std::unique_lock<std::mutex> lk(m_mutex);
                m_sleepers.fetch_add(1);
                m_cv.wait(lk, [&] { return m_generation.load(std::memory_order_acquire) != seen; });
                m_sleepers.fetch_sub(1);
            }
        }
        seen = m_generation.load(std::memory_order_acquire);
        if (m_stop.load()) {
            return;
        }
        (*m_job)(tid);
        m_done.fetch_add(1, std::memory_order_release);
    }
}

void BatchPlanner::Pool::run(const std::function<void(int)>& f)
{
    if (m_n == 1) {
        f(0);
        return;
    }
    m_job = &f;
    m_done.store(0, std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lk(m_mutex);   // a worker between its predicate check and its wait cannot miss this
        m_generation.fetch_add(1, std::memory_order_release);
    }
    if (m_sleepers.load() > 0) {
        m_cv.notify_all();
    }
    f(0);
    int spins = 0;
    while (m_done.load(std::memory_order_acquire) != m_n - 1) {
        if (++spins > 256) {
            std::this_thread::yield();
        }
    }
}

///////////////////////////////////////////////////////////////////////////////
// per-query steps
///////////////////////////////////////////////////////////////////////////////

void BatchPlanner::initQuery(Query& Q, int index, int slot, const double* goal)
{
    Q = Query();
    Q.index = index;
    Q.slot = slot;
    Q.done = false;
    Q.expanding = -1;
    Q.n_succ = 0;
    Q.edge_begin = 0;
    for (int a = 0; a < 3; ++a) Q.goal[a] = goal[a];
    int cell[3];
    worldToGrid(Q.goal, cell);
    const bool inb = cell[0] >= 0 && cell[1] >= 0 && cell[2] >= 0 &&
                     cell[0] < m_cfg.dims[0] && cell[1] < m_cfg.dims[1] && cell[2] < m_cfg.dims[2];
    // the seed cell holds distance 0 even if it was a wall (bfs3d.cpp:181-187)
    Q.goal_h = inb ? 0 : SMPLGPU_HEURISTIC_INFINITY;
    Q.open.assign(1, -1);
    Q.lat.dof = m_cfg.dof;
    Q.lat.add(nullptr, nullptr, 0, 0, false); // id 0 = the goal state (manip_lattice.cpp:122)
}

// ARAStar::improvePath loop head (arastar.cpp:486-527) + ManipLatticeActionSpace::apply + the joint-limit
// part of ManipLattice::checkAction
void BatchPlanner::expandOne(Query& Q)
{
    const int dof = m_cfg.dof;
    Q.expanding = -1;
    Q.n_succ = 0;
    Q.succ_q1.clear();
    if (Q.open.size() <= 1) {
        finish(Q, false);
        return;
    }
    const int min_id = Q.open[1];
    if (Q.search[min_id].f >= Q.search[0].f || min_id == 0) {
        finish(Q, true);
        return;
    }
    if (Q.result.expansions >= m_cfg.max_expansions) {
        finish(Q, false);
        return;
    }
    heapPop(Q);
    Q.search[min_id].iteration_closed = 1;
    Q.search[min_id].eg = Q.search[min_id].g;
    Q.expanding = min_id;
    ++Q.result.expansions;

    const double* Pq = Q.lat.q(min_id);
    const double goal_dist = (double)Q.lat.gdist[min_id] * m_cfg.res;
    const bool near_goal = goal_dist <= m_cfg.short_dist_thresh;
    for (size_t p = 0; p < m_prim_deltas.size(); ++p) {
        const bool active_prim = m_prim_short[p] ? (m_cfg.use_short_dist && near_goal)
                                                 : !(m_cfg.use_short_dist && near_goal);
        if (!active_prim) {
            continue;
        }
        const size_t base = Q.succ_q1.size();
        Q.succ_q1.resize(base + dof);
        for (int j = 0; j < dof; ++j) {
            Q.succ_q1[base + j] = m_prim_deltas[p][j] + Pq[j];
        }
        if (!checkJointLimits(&Q.succ_q1[base])) { // checkAction: joint limits first
            Q.succ_q1.resize(base);
            continue;
        }
        ++Q.n_succ;
    }
}

// GetSuccs bookkeeping + ARAStar::expand relaxations (arastar.cpp:531-568) for this query's edges, in
// submission (= primitive) order
void BatchPlanner::absorbOne(Query& Q, const uint8_t* verdict, const int32_t* h, const int32_t* gd, const double* off)
{
    const int dof = m_cfg.dof;
    std::vector<int> coord;
    for (int e = 0; e < Q.n_succ; ++e) {
        if (!verdict[e]) {
            continue;
        }
        const double* qs = &Q.succ_q1[(size_t)e * dof];
        stateToCoord(qs, coord);
        int succ_id = Q.lat.find(coord.data());
        if (succ_id < 0) {
            succ_id = Q.lat.add(coord.data(), qs, h[e], gd[e], true);
        }
        // ManipLattice::isGoal, XYZ_GOAL (manip_lattice.cpp:1673-1687)
        const bool is_goal = std::fabs(off[3 * e] - Q.goal[0]) <= m_cfg.xyz_tolerance[0] &&
                             std::fabs(off[3 * e + 1] - Q.goal[1]) <= m_cfg.xyz_tolerance[1] &&
                             std::fabs(off[3 * e + 2] - Q.goal[2]) <= m_cfg.xyz_tolerance[2];
        if (is_goal && (Q.goal_succ.empty() || Q.goal_succ.back().first != Q.expanding)) {
            Q.goal_succ.emplace_back(Q.expanding, succ_id);   // first valid goal action of this expansion
        }
        const int target = is_goal ? 0 : succ_id;
        sstate(Q, target);
        touch(Q, target);
        SState& ss = Q.search[target];
        const int new_cost = Q.search[Q.expanding].eg + (int)(1000 * 1.0);
        if ((unsigned int)new_cost < ss.g) {   // int vs unsigned, compared as unsigned (arastar.cpp:545-548)
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
}

///////////////////////////////////////////////////////////////////////////////
// the lock-step driver
///////////////////////////////////////////////////////////////////////////////

bool BatchPlanner::plan(const double* starts, const double* goals, int nq, std::vector<QueryResult>& out, std::string* err)
{
    out.assign(nq, QueryResult());
    m_stats = BatchStats();
    if (nq == 0) {
        return true;
    }
    const int dof = m_cfg.dof;
    Timer t;
    auto fail_dev = [&]() {
        if (err) *err = smplgpu_last_error(m_ctx);
        return false;
    };
    const long long resolved0 = smplgpu_expand_batch_resolved(m_ctx);
    // the bank keeps the reference's int node indices and has to fit in device memory: clamp the concurrency to it
    const int max_slots = smplgpu_bfs_bank_max_slots(m_ctx);
    if (max_slots < 0) return fail_dev();
    const int n_slots = std::min(std::min(m_max_concurrent, nq), max_slots);
    if (smplgpu_bfs_bank_create(m_ctx, n_slots, m_cfg.inflation_radius) < 0) return fail_dev();
    // every device / pinned allocation happens here: allocating while other planner threads run would stall
    // their streams (allocation synchronises the device)
    if (smplgpu_expand_batch_reserve(m_ctx, (n_slots + 1) * (int)m_prim_deltas.size()) < 0) return fail_dev();
    ++m_stats.device_calls;
    {
        const double dt = t.lap();
        m_stats.device_seconds += dt;
        m_stats.setup_seconds += dt;
    }

    Pool pool(m_cfg.n_threads);
    std::vector<Query> S(n_slots);          // slot -> query occupying it
    std::vector<char> occupied(n_slots, 0);
    int next_query = 0, finished = 0;
    // refill in groups so that one bank run (one wavefront launch sequence) serves several new queries
    const int refill_min = std::max(1, n_slots / 4);

    // The active queries work in two groups (slot parity).  While the device expands one group's batch the
    // host absorbs the other group's results and prepares its next batch, so host and device overlap.
    struct Group
    {
        std::vector<int> active;            // occupied slots of this group with a search in progress
        std::vector<double> q0, q1, off;
        std::vector<int32_t> slot, h, gd;
        std::vector<uint8_t> verdict;
        bool in_flight = false;
        int ne = 0;
    };
    Group G[2];
    std::vector<double> sq;                 // setStart scratch
    std::vector<int32_t> sh, sgd;
    std::vector<uint8_t> sv;
    std::vector<double> soff;

    // wait for a group's batch, relax its queries, retire the finished ones
    auto absorb = [&](int gi) -> bool {
        Group& g = G[gi];
        if (g.in_flight) {
            m_stats.host_seconds += t.lap();
            if (smplgpu_expand_batch_wait(m_ctx, gi, g.verdict.data(), g.h.data(), g.gd.data(), g.off.data()) < 0) return false;
            {
                const double dt = t.lap();
                m_stats.device_seconds += dt;
                m_stats.max_wait_seconds = std::max(m_stats.max_wait_seconds, dt);
            }
            g.in_flight = false;
            const int na = (int)g.active.size();
            pool.run([&](int tid) {
                for (int k = tid; k < na; k += pool.size()) {
                    Query& Q = S[g.active[k]];
                    if (Q.done || Q.n_succ == 0) {
                        continue;
                    }
                    absorbOne(Q, g.verdict.data() + Q.edge_begin, g.h.data() + Q.edge_begin, g.gd.data() + Q.edge_begin,
                              g.off.data() + (size_t)Q.edge_begin * 3);
                }
            });
        }
        size_t keep = 0;
        for (size_t k = 0; k < g.active.size(); ++k) {
            const int s = g.active[k];
            if (S[s].done) {
                out[S[s].index] = S[s].result;
                occupied[s] = 0;
                ++finished;
            } else {
                g.active[keep++] = s;
            }
        }
        g.active.resize(keep);
        return true;
    };

    // every active query of the group pops one state and generates its successors; one device batch
    auto expand = [&](int gi) -> bool {
        Group& g = G[gi];
        const int na = (int)g.active.size();
        if (na == 0) {
            return true;
        }
        pool.run([&](int tid) {
            for (int k = tid; k < na; k += pool.size()) {
                Query& Q = S[g.active[k]];
                if (!Q.done) {
                    expandOne(Q);
                }
            }
        });
        int ne = 0;
        for (int k = 0; k < na; ++k) {
            Query& Q = S[g.active[k]];
            Q.edge_begin = ne;
            ne += Q.done ? 0 : Q.n_succ;
        }
        g.ne = ne;
        if (ne == 0) {
            return true;   // finished / successor-less queries are retired by the next absorb
        }
        g.q0.resize((size_t)ne * dof);
        g.q1.resize((size_t)ne * dof);
        g.slot.resize(ne);
        pool.run([&](int tid) {
            for (int k = tid; k < na; k += pool.size()) {
                const Query& Q = S[g.active[k]];
                if (Q.done || Q.n_succ == 0) {
                    continue;
                }
                const double* pq = Q.lat.q(Q.expanding);
                for (int e = 0; e < Q.n_succ; ++e) {
                    std::copy(pq, pq + dof, g.q0.begin() + (size_t)(Q.edge_begin + e) * dof);
                    g.slot[Q.edge_begin + e] = Q.slot;
                }
                std::copy(Q.succ_q1.begin(), Q.succ_q1.end(), g.q1.begin() + (size_t)Q.edge_begin * dof);
            }
        });
        g.verdict.resize(ne);
        g.h.resize(ne);
        g.gd.resize(ne);
        g.off.resize((size_t)ne * 3);
        if (smplgpu_expand_batch_submit(m_ctx, g.q0.data(), g.q1.data(), g.slot.data(), ne, m_cfg.cost_per_cell, gi) < 0) return false;
        g.in_flight = true;
        ++m_stats.device_calls;
        ++m_stats.rounds;
        m_stats.edges_submitted += ne;
        return true;
    };

    while (finished < nq) {
        // ---- hand free slots to waiting queries: setGoal (one BFS per query) + setStart ----
        int n_free = 0;
        for (int s = 0; s < n_slots; ++s) n_free += occupied[s] ? 0 : 1;
        const bool idle = G[0].active.empty() && G[1].active.empty();
        if (next_query < nq && (n_free >= refill_min || idle)) {
            // the refill uses the synchronous entry points: nothing may be in flight
            for (int gi = 0; gi < 2; ++gi) {
                if (!absorb(gi)) return fail_dev();
            }
            std::vector<int32_t> new_slots, seeds;
            for (int s = 0; s < n_slots && next_query < nq; ++s) {
                if (occupied[s]) {
                    continue;
                }
                initQuery(S[s], next_query, s, goals + (size_t)next_query * 3);
                ++next_query;
                occupied[s] = 1;
                new_slots.push_back(s);
                int cell[3];
                worldToGrid(S[s].goal, cell);
                seeds.insert(seeds.end(), cell, cell + 3);
            }
            const int nn = (int)new_slots.size();
            m_stats.host_seconds += t.lap();
            if (smplgpu_bfs_bank_run_slots(m_ctx, new_slots.data(), seeds.data(), nn) < 0) return fail_dev();
            ++m_stats.bfs_runs;
            // setStart: limits + validity, then heuristic / metric distance of the start state
            sq.resize((size_t)nn * dof);
            for (int k = 0; k < nn; ++k) {
                const int qi = S[new_slots[k]].index;
                std::copy(starts + (size_t)qi * dof, starts + (size_t)(qi + 1) * dof, sq.begin() + (size_t)k * dof);
            }
            sv.resize(nn);
            sh.resize(nn);
            sgd.resize(nn);
            soff.resize((size_t)nn * 3);
            std::vector<uint8_t> dummy(nn);
            if (smplgpu_is_states_valid(m_ctx, sq.data(), nn, sv.data()) < 0) return fail_dev();
            if (smplgpu_expand_batch(m_ctx, sq.data(), sq.data(), new_slots.data(), nn, m_cfg.cost_per_cell,
                                     dummy.data(), sh.data(), sgd.data(), soff.data()) < 0) return fail_dev();
            m_stats.device_calls += 3;
            {
                const double dt = t.lap();
                m_stats.device_seconds += dt;
                m_stats.setup_seconds += dt;
            }
            std::vector<int> coord;
            for (int k = 0; k < nn; ++k) {
                Query& Q = S[new_slots[k]];
                const double* qs = &sq[(size_t)k * dof];
                G[new_slots[k] & 1].active.push_back(new_slots[k]);
                if (!checkJointLimits(qs) || !sv[k]) {
                    finish(Q, false);   // retired by the group's next absorb
                    continue;
                }
                stateToCoord(qs, coord);
                Q.lat.add(coord.data(), qs, sh[k], sgd[k], true);
                sstate(Q, 1);
                touch(Q, 1);
                touch(Q, 0);
                Q.search[1].g = 0;
                Q.search[1].f = computeKey(Q.search[1]);
                heapPush(Q, 1);
            }
        }

        // ---- one pipelined round: each group absorbs its previous batch and submits the next ----
        for (int gi = 0; gi < 2; ++gi) {
            if (!absorb(gi)) return fail_dev();
            if (!expand(gi)) return fail_dev();
        }
        m_stats.host_seconds += t.lap();
    }
    m_stats.edges_resolved_f64 = smplgpu_expand_batch_resolved(m_ctx) - resolved0;
    // nothing can be in flight here: a group's queries finish in absorb or expand, and a batch is only
    // submitted for queries that are not done; drain defensively anyway
    for (int gi = 0; gi < 2; ++gi) {
        if (G[gi].in_flight) {
            if (smplgpu_expand_batch_wait(m_ctx, gi, G[gi].verdict.data(), G[gi].h.data(), G[gi].gd.data(), G[gi].off.data()) < 0) return fail_dev();
        }
    }
    return true;
}

} // namespace smplhost
