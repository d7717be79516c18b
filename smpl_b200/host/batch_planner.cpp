#include "batch_planner.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
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

// A table slot holds (low 32 hash bits << 32 | state id): a probe is decided from the slot alone unless the tag
// matches, so the coordinate array (another cache miss) is only read for a state that is almost surely the one
int BatchPlanner::Lattice::find(const int* c, uint64_t hv) const
{
    if (table.empty()) {
        return -1;
    }
    const size_t mask = table.size() - 1;
    const uint32_t tag = (uint32_t)hv;
    for (size_t i = (size_t)tag & mask;; i = (i + 1) & mask) {
        const uint64_t slot = table[i];
        if (slot == EMPTY) {
            return -1;
        }
        if ((uint32_t)(slot >> 32) == tag) {
            const int id = (int)(uint32_t)slot;
            if (std::equal(c, c + dof, &coords[(size_t)id * dof])) {
                return id;
            }
        }
    }
}

void BatchPlanner::Lattice::enter(uint32_t tag, int id)
{
    const size_t mask = table.size() - 1;
    size_t i = (size_t)tag & mask;
    while (table[i] != EMPTY) {
        i = (i + 1) & mask;
    }
    table[i] = ((uint64_t)tag << 32) | (uint32_t)id;
}

void BatchPlanner::Lattice::grow(uint32_t new_tag, int new_id)
{
    // the tags carry every hash bit a table of up to 2^32 slots needs: re-enter the old slots without reading
    // a single coordinate, then the state that triggered the growth
    std::vector<uint64_t> old;
    old.swap(table);
    table.assign(old.empty() ? 256 : old.size() * 2, EMPTY);
    for (const uint64_t slot : old) {
        if (slot != EMPTY) {
            enter((uint32_t)(slot >> 32), (int)(uint32_t)slot);
        }
    }
    enter(new_tag, new_id);
}

int BatchPlanner::Lattice::add(const int* c, const double* q, int hval, int gd, bool index, uint64_t hv)
{
    const int id = size();
    if (c != nullptr) {
        coords.insert(coords.end(), c, c + dof);
        qs.insert(qs.end(), q, q + dof);
    } else {
        coords.insert(coords.end(), dof, 0);
        qs.insert(qs.end(), dof, 0.0);
    }
    h.push_back(hval);
    gdist.push_back(gd);
    if (index) {
        if ((size_t)(id + 1) * 2 > table.size()) {
            grow((uint32_t)hv, id);
        } else {
            enter((uint32_t)hv, id);
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
    const Query::HeapEntry tmp = Q.open[pivot];
    while (pivot != 1) {
        const size_t p = pivot >> 1;
        if (Q.open[p].f < tmp.f) {
            break;
        }
        Q.open[pivot] = Q.open[p];
        Q.search[Q.open[pivot].id].heap_index = (int)pivot;
        pivot = p;
    }
    Q.open[pivot] = tmp;
    Q.search[tmp.id].heap_index = (int)pivot;
}

void BatchPlanner::percolateDown(Query& Q, size_t pivot)
{
    if (pivot >= Q.open.size()) {
        return;
    }
    size_t left = pivot << 1, right = left + 1;
    const Query::HeapEntry tmp = Q.open[pivot];
    while (left < Q.open.size()) {
        size_t s = right;
        if (right >= Q.open.size() || Q.open[left].f < Q.open[right].f) {
            s = left;
        }
        if (Q.open[s].f < tmp.f) {
            Q.open[pivot] = Q.open[s];
            Q.search[Q.open[pivot].id].heap_index = (int)pivot;
            pivot = s;
        } else {
            break;
        }
        left = pivot << 1;
        right = left + 1;
    }
    Q.open[pivot] = tmp;
    Q.search[tmp.id].heap_index = (int)pivot;
}

void BatchPlanner::heapPush(Query& Q, int id)
{
    Q.search[id].heap_index = (int)Q.open.size();
    Q.open.push_back(Query::HeapEntry{ Q.search[id].f, id });
    percolateUp(Q, Q.open.size() - 1);
}

void BatchPlanner::heapPop(Query& Q)
{
    Q.search[Q.open[1].id].heap_index = 0;
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
    // keep the slot's allocations from the query it served before (growing a multi-megabyte vector means mmap +
    // copy + munmap, and every munmap interrupts all planner threads); a fresh slot reserves room up front
    Query fresh = Query();
    std::swap(fresh.lat.coords, Q.lat.coords);
    std::swap(fresh.lat.qs, Q.lat.qs);
    std::swap(fresh.lat.h, Q.lat.h);
    std::swap(fresh.lat.gdist, Q.lat.gdist);
    std::swap(fresh.lat.table, Q.lat.table);
    std::swap(fresh.search, Q.search);
    std::swap(fresh.open, Q.open);
    std::swap(fresh.succ_q1, Q.succ_q1);
    std::swap(fresh.succ_coord, Q.succ_coord);
    std::swap(fresh.succ_hslot, Q.succ_hslot);
    std::swap(fresh.succ_id, Q.succ_id);
    fresh.lat.coords.clear();
    fresh.lat.qs.clear();
    fresh.lat.h.clear();
    fresh.lat.gdist.clear();
    fresh.lat.table.clear();
    fresh.search.clear();
    fresh.open.clear();
    Q = std::move(fresh);
    {
        const size_t per_expansion = std::max<size_t>(1, m_prim_deltas.size() / 2);
        const size_t n = std::min<size_t>(2 + (size_t)std::max(0, m_cfg.max_expansions) * per_expansion, (size_t)1 << 15);
        const size_t dofs = (size_t)m_cfg.dof;
        Q.lat.coords.reserve(n * dofs);
        Q.lat.qs.reserve(n * dofs);
        Q.lat.h.reserve(n);
        Q.lat.gdist.reserve(n);
        Q.search.reserve(n);
    }
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
    Q.open.assign(1, Query::HeapEntry{ 0u, -1 });
    Q.lat.dof = m_cfg.dof;
    Q.lat.add(nullptr, nullptr, 0, 0, false, 0); // id 0 = the goal state (manip_lattice.cpp:122)
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
    const int min_id = Q.open[1].id;
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
static thread_local double g_abs_prof[4] = { 0, 0, 0, 0 };   // SMPLHOST_PLAN_PROFILE: coord+hash, prefetch pass, lookup, relax

void BatchPlanner::absorbOne(Query& Q, const uint8_t* verdict, const int32_t* h, const int32_t* gd, const double* off)
{
    const int dof = m_cfg.dof;
    static const bool profile = getenv("SMPLHOST_PLAN_PROFILE") != nullptr;
    Timer pt;
    std::vector<int> coord;
    // The lattice of a query is megabytes of hash table, coordinates and search states, and hundreds of queries
    // take turns on one thread: every lookup below misses the caches.  Two prefetch-only passes (they decide
    // nothing) start those loads for all successors of this expansion before the sequential relaxation runs.
    Q.succ_coord.resize((size_t)Q.n_succ * dof);
    Q.succ_hslot.resize(Q.n_succ);
    const size_t mask = Q.lat.table.empty() ? 0 : Q.lat.table.size() - 1;
    for (int e = 0; e < Q.n_succ; ++e) {
        if (!verdict[e]) {
            continue;
        }
        stateToCoord(&Q.succ_q1[(size_t)e * dof], coord);
        std::copy(coord.begin(), coord.end(), Q.succ_coord.begin() + (size_t)e * dof);
        Q.succ_hslot[e] = Lattice::hash(coord.data(), dof);
        if (!Q.lat.table.empty()) {
            __builtin_prefetch(&Q.lat.table[(size_t)Q.succ_hslot[e] & mask]);
        }
    }
    if (profile) g_abs_prof[0] += pt.lap();
    for (int e = 0; e < Q.n_succ; ++e) {
        if (!verdict[e] || Q.lat.table.empty()) {
            continue;
        }
        const uint64_t slot = Q.lat.table[(size_t)Q.succ_hslot[e] & mask];
        const int id = slot == Lattice::EMPTY ? -1 : (int)(uint32_t)slot;
        if (id >= 0) {
            __builtin_prefetch(&Q.lat.coords[(size_t)id * dof]);
            if ((size_t)id < Q.search.size()) {
                __builtin_prefetch(&Q.search[id]);
            }
        }
    }
    if (profile) g_abs_prof[1] += pt.lap();
    Q.succ_id.resize(Q.n_succ);
    int* ids = Q.succ_id.data();   // lattice id of successor e
    for (int e = 0; e < Q.n_succ; ++e) {
        if (!verdict[e]) {
            continue;
        }
        const int* c = &Q.succ_coord[(size_t)e * dof];
        int succ_id = Q.lat.find(c, Q.succ_hslot[e]);
        if (succ_id < 0) {
            succ_id = Q.lat.add(c, &Q.succ_q1[(size_t)e * dof], h[e], gd[e], true, Q.succ_hslot[e]);
        }
        ids[e] = succ_id;
    }
    if (profile) g_abs_prof[2] += pt.lap();
    for (int e = 0; e < Q.n_succ; ++e) {
        if (!verdict[e]) {
            continue;
        }
        const int succ_id = ids[e];
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
                    Q.open[ss.heap_index].f = ss.f;
                    percolateUp(Q, (size_t)ss.heap_index);
                } else {
                    heapPush(Q, target);
                }
            }
        }
    }
    // the state this query expands next round (unless a better one arrives): start loading what expandOne reads
    if (Q.open.size() > 1) {
        const int next_id = Q.open[1].id;
        __builtin_prefetch(&Q.search[next_id]);
        __builtin_prefetch(Q.lat.q(next_id));
        __builtin_prefetch(&Q.lat.gdist[next_id]);
    }
    if (profile) g_abs_prof[3] += pt.lap();
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

    // The active queries work in NG groups (slot modulo NG).  While the device expands one group's batch the
    // host absorbs the other groups' results and prepares their next batches, so host and device overlap.  Two
    // groups is the measured optimum (SMPLHOST_PLAN_GROUPS, 2 .. SMPLGPU_EXPAND_BUFFERS: 12 planner threads on one
    // B200 reach 1651 queries/s with 2 groups, 1346 with 3, 1056 with 4): the host does wait for the device a fifth
    // of the time, but more groups mean more and smaller batches, and the device side of a batch is a fixed-latency
    // chain (copy in, three kernels, copy out) that twelve contexts already queue behind one another.
    struct Group
    {
        std::vector<int> active;            // occupied slots of this group with a search in progress
        std::vector<double> q0, q1, off;
        std::vector<int32_t> slot, h, gd;
        std::vector<uint8_t> verdict;
        bool in_flight = false;
        int ne = 0;
    };
    static const int NG = [] {
        const char* e = getenv("SMPLHOST_PLAN_GROUPS");
        const int v = e ? atoi(e) : 2;
        return std::min(std::max(v, 2), (int)SMPLGPU_EXPAND_BUFFERS);
    }();
    std::vector<Group> G(NG);
    auto all_idle = [&]() {
        for (const Group& g : G) {
            if (!g.active.empty()) return false;
        }
        return true;
    };
    // SMPLHOST_PLAN_PROFILE=1: where the host time of this planner thread goes (printed to stderr at the end)
    static const bool profile = getenv("SMPLHOST_PLAN_PROFILE") != nullptr;
    double prof[6] = { 0, 0, 0, 0, 0, 0 };  // absorb, retire, expand, pack, submit, refill
    Timer pt;
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
            pt.lap();
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
            prof[0] += pt.lap();
        }
        pt.lap();
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
        prof[1] += pt.lap();
        return true;
    };

    // every active query of the group pops one state and generates its successors; one device batch
    auto expand = [&](int gi) -> bool {
        Group& g = G[gi];
        const int na = (int)g.active.size();
        if (na == 0) {
            return true;
        }
        pt.lap();
        pool.run([&](int tid) {
            for (int k = tid; k < na; k += pool.size()) {
                Query& Q = S[g.active[k]];
                if (!Q.done) {
                    expandOne(Q);
                }
            }
        });
        prof[2] += pt.lap();
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
        prof[3] += pt.lap();
        if (smplgpu_expand_batch_submit(m_ctx, g.q0.data(), g.q1.data(), g.slot.data(), ne, m_cfg.cost_per_cell, gi) < 0) return false;
        prof[4] += pt.lap();
        g.in_flight = true;
        ++m_stats.device_calls;
        ++m_stats.rounds;
        m_stats.edges_submitted += ne;
        return true;
    };

    // Admission in two steps: (1) setGoal -- a free slot gets the query's goal and its BFS is queued on the device
    // WITHOUT waiting (the reference's BFS_3D also searches in a background thread, bfs3d.cpp:156-201), the running
    // searches keep expanding; (2) when that BFS is done, setStart (limits, validity, heuristic of the start state)
    // and the queries join the rounds.  One group of pending slots at a time.
    std::vector<int32_t> pending;           // slots whose BFS is in flight
    std::vector<int32_t> seeds;
    bool first_fill = true;

    auto activate = [&]() -> bool {
        // setStart uses the synchronous entry points: nothing of ours may be in flight
        for (int gi = 0; gi < NG; ++gi) {
            if (!absorb(gi)) return false;
        }
        m_stats.host_seconds += t.lap();
        if (smplgpu_bfs_bank_run_wait(m_ctx) < 0) return false;
        const int nn = (int)pending.size();
        sq.resize((size_t)nn * dof);
        for (int k = 0; k < nn; ++k) {
            const int qi = S[pending[k]].index;
            std::copy(starts + (size_t)qi * dof, starts + (size_t)(qi + 1) * dof, sq.begin() + (size_t)k * dof);
        }
        sv.resize(nn);
        sh.resize(nn);
        sgd.resize(nn);
        soff.resize((size_t)nn * 3);
        std::vector<uint8_t> dummy(nn);
        if (smplgpu_is_states_valid(m_ctx, sq.data(), nn, sv.data()) < 0) return false;
        if (smplgpu_expand_batch(m_ctx, sq.data(), sq.data(), pending.data(), nn, m_cfg.cost_per_cell,
                                 dummy.data(), sh.data(), sgd.data(), soff.data()) < 0) return false;
        m_stats.device_calls += 2;
        {
            const double dt = t.lap();
            m_stats.device_seconds += dt;
            m_stats.setup_seconds += dt;
        }
        std::vector<int> coord;
        for (int k = 0; k < nn; ++k) {
            Query& Q = S[pending[k]];
            const double* qs = &sq[(size_t)k * dof];
            G[pending[k] % NG].active.push_back(pending[k]);
            if (!checkJointLimits(qs) || !sv[k]) {
                finish(Q, false);   // retired by the group's next absorb
                continue;
            }
            stateToCoord(qs, coord);
            Q.lat.add(coord.data(), qs, sh[k], sgd[k], true, Lattice::hash(coord.data(), dof));
            sstate(Q, 1);
            touch(Q, 1);
            touch(Q, 0);
            Q.search[1].g = 0;
            Q.search[1].f = computeKey(Q.search[1]);
            heapPush(Q, 1);
        }
        pending.clear();
        return true;
    };

    while (finished < nq) {
        const bool idle = all_idle();
        // ---- (2) the pending queries' BFS is done (or there is nothing else to do): setStart, join the rounds ----
        if (!pending.empty() && (idle || smplgpu_bfs_bank_run_done(m_ctx) != 0)) {
            if (!activate()) return fail_dev();
            continue;
        }
        // ---- (1) hand free slots to waiting queries: setGoal (one BFS per query, queued) ----
        int n_free = 0;
        for (int s = 0; s < n_slots; ++s) n_free += occupied[s] ? 0 : 1;
        if (pending.empty() && next_query < nq && (n_free >= refill_min || idle)) {
            // the first group is kept small so that the searches start while the other slots' BFS still runs
            // (measured on one B200, 12 contexts x 171 slots: first group = 1/2 of the slots 1729 queries/s, 1/4 1655,
            // 1/8 at every admission 1425 -- every bank run pays whole-bank passes; SMPLHOST_ADMIT_CHUNKS overrides)
            static const int chunks = getenv("SMPLHOST_ADMIT_CHUNKS") ? std::max(1, atoi(getenv("SMPLHOST_ADMIT_CHUNKS"))) : 2;
            const int take = (first_fill || chunks > 4) ? std::max(1, n_slots / chunks) : n_slots;
            first_fill = false;
            seeds.clear();
            for (int s = 0; s < n_slots && next_query < nq && (int)pending.size() < take; ++s) {
                if (occupied[s]) {
                    continue;
                }
                initQuery(S[s], next_query, s, goals + (size_t)next_query * 3);
                ++next_query;
                occupied[s] = 1;
                pending.push_back(s);
                int cell[3];
                worldToGrid(S[s].goal, cell);
                seeds.insert(seeds.end(), cell, cell + 3);
            }
            m_stats.host_seconds += t.lap();
            if (smplgpu_bfs_bank_run_slots_async(m_ctx, pending.data(), seeds.data(), (int)pending.size()) < 0) return fail_dev();
            ++m_stats.bfs_runs;
            ++m_stats.device_calls;
            {
                const double dt = t.lap();
                m_stats.device_seconds += dt;
                m_stats.setup_seconds += dt;
            }
            if (all_idle()) {
                continue;   // nothing to expand meanwhile: go and wait for it
            }
        }

        // ---- one pipelined round: each group absorbs its previous batch and submits the next ----
        for (int gi = 0; gi < NG; ++gi) {
            if (!absorb(gi)) return fail_dev();
            if (!expand(gi)) return fail_dev();
        }
        m_stats.host_seconds += t.lap();
    }
    if (profile) {
        fprintf(stderr, "[plan profile] rounds %d edges %lld expansions/ctx: absorb %.3f retire %.3f expand %.3f pack %.3f submit %.3f | "
                        "device wait %.3f setup %.3f host %.3f s\n", m_stats.rounds, m_stats.edges_submitted, prof[0], prof[1],
                prof[2], prof[3], prof[4], m_stats.device_seconds - m_stats.setup_seconds, m_stats.setup_seconds,
                m_stats.host_seconds);
        fprintf(stderr, "[plan profile]   absorb: coord+hash %.3f prefetch %.3f lookup/insert %.3f relax/heap %.3f s\n",
                g_abs_prof[0], g_abs_prof[1], g_abs_prof[2], g_abs_prof[3]);
        for (double& v : g_abs_prof) v = 0.0;
    }
    m_stats.edges_resolved_f64 = smplgpu_expand_batch_resolved(m_ctx) - resolved0;
    // nothing can be in flight here: a group's queries finish in absorb or expand, and a batch is only
    // submitted for queries that are not done; drain defensively anyway
    for (int gi = 0; gi < NG; ++gi) {
        if (G[gi].in_flight) {
            if (smplgpu_expand_batch_wait(m_ctx, gi, G[gi].verdict.data(), G[gi].h.data(), G[gi].gd.data(), G[gi].off.data()) < 0) return fail_dev();
        }
    }
    return true;
}

} // namespace smplhost
