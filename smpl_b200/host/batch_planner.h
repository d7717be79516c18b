// Host side of the caller of the hot path, B200-first: MANY independent ARA* searches on the
// manipulation lattice advance in lock step, and every round submits the successors of all of
// them to the device in ONE smplgpu_expand_batch call (edge validity + heuristic + goal test
// inputs), so the per-launch batch is (active queries x ~8-22 edges) instead of one edge.
//
// It mirrors, per query, the reference's
//   ManipLattice::GetSuccs / checkAction / isGoal / stateToCoord   smpl/src/graph/manip_lattice.cpp:219-313, 1263-1289, 1511-1580, 1673-1687
//   ManipLatticeActionSpace::apply / mprimActive                   smpl/src/graph/manip_lattice_action_space.cpp:376-449, 662-691
//   ARAStar::replan / improvePath / expand / computeKey            smpl/src/search/arastar.cpp:107-215, 486-568, 579-582
//   intrusive_heap                                                 smpl/include/smpl/detail/intrusive_heap.hpp
// so that each query returns the same path, cost and expansion count as the reference-shaped
// sequential planner (first solution at the initial epsilon; see SURVEY.md section 8 defect 2 for
// the motion-primitive format and goal type decisions).  The OPEN lists, hash tables and the
// lattice stay on the host, as BASELINE.json's north_star prescribes.
#ifndef SMPLHOST_BATCH_PLANNER_H
#define SMPLHOST_BATCH_PLANNER_H

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <cstdint>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/smplgpu.h"

namespace smplhost {

struct PlannerConfig
{
    int dof = 0;
    std::vector<double> resolutions;
    std::vector<double> mprims;         // n_prims x dof deltas, file order (converses are added here)
    std::vector<uint8_t> short_flags;   // n_prims
    bool use_short_dist = true;
    double short_dist_thresh = 0.4;
    double epsilon = 100.0;
    int max_expansions = 200000;
    double xyz_tolerance[3] = { 0.015, 0.015, 0.015 };
    int cost_per_cell = 100;
    double inflation_radius = 0.02;
    // joint limits as KDLRobotModel reports them (min/max/continuous per planning variable)
    std::vector<double> var_min, var_max;
    std::vector<int> var_continuous;
    // grid geometry for host-side worldToGrid of the goal
    double origin[3] = { 0, 0, 0 };
    double res = 0.02;
    int dims[3] = { 0, 0, 0 };
    // host threads for the per-query work (successor generation, hashing, OPEN list updates); queries are
    // independent, so the result does not depend on this
    int n_threads = 1;
};

struct QueryResult
{
    bool success = false;
    int expansions = 0;
    int cost = 0;
    int num_states = 0;
    std::vector<int> path_ids;
    // ManipLattice::extractPath (manip_lattice.cpp:2018-2160): the joint-space state of every path id; the goal id
    // is replaced by the lattice state of the first valid goal-reaching action of its predecessor
    std::vector<double> path_states;   // path_ids.size() x dof
};

struct BatchStats
{
    int rounds = 0;
    int bfs_runs = 0;              // bank runs (each covers every slot that was refilled at that point)
    long long edges_submitted = 0;
    long long edges_resolved_f64 = 0;   // of those, decided by the double-precision kernels
    long long device_calls = 0;
    double device_seconds = 0.0;   // time inside smplgpu_* calls
    double host_seconds = 0.0;     // everything else
    double setup_seconds = 0.0;    // of device_seconds: bank creation, BFS bank runs, setStart (per refill)
    double max_wait_seconds = 0.0; // longest single wait for an expansion batch
};

class BatchPlanner
{
public:
    BatchPlanner(smplgpu_ctx* ctx, const PlannerConfig& cfg, int max_concurrent);

    /// starts: nq x dof, goals: nq x 3 (target-offset position in the planning frame)
    bool plan(const double* starts, const double* goals, int nq, std::vector<QueryResult>& out, std::string* err);

    const BatchStats& stats() const { return m_stats; }

private:
    // Lattice states of one query, flat (no per-state allocation): state id -> coord[dof], q[dof], h, gdist,
    // and an open-addressing hash table coord -> id (the role of ManipLattice's m_state_to_id, manip_lattice.h)
    struct Lattice
    {
        int dof = 0;
        std::vector<int> coords;
        std::vector<double> qs;
        std::vector<int> h, gdist;
        std::vector<uint64_t> table; // size is a power of two; slot = low 32 hash bits << 32 | state id, EMPTY = all ones
        static constexpr uint64_t EMPTY = ~0ull;
        int size() const { return (int)h.size(); }
        const double* q(int id) const { return &qs[(size_t)id * dof]; }
        static uint64_t hash(const int* c, int dof)
        {
            uint64_t x = 0x9E3779B97F4A7C15ull;
            for (int i = 0; i < dof; ++i) {
                x ^= (uint64_t)(uint32_t)c[i] + 0x9E3779B97F4A7C15ull + (x << 6) + (x >> 2);
            }
            return x;
        }
        int find(const int* c, uint64_t hv) const;
        int add(const int* c, const double* q, int hval, int gd, bool index, uint64_t hv);   // index = enter it in the table
        void grow(uint32_t new_tag, int new_id);
        void enter(uint32_t tag, int id);
    };
    // g, h, f, eg are unsigned as in the reference's ARAStar::SearchState (arastar.h:176-187): a negative heuristic
    // (unreachable BFS cell: cost_per_cell * -1) sorts last in OPEN
    struct SState { unsigned int g, h, f, eg; int iteration_closed, bp, heap_index; bool touched; };
    struct Query
    {
        int index;
        int slot;
        double goal[3];
        int goal_h;
        Lattice lat;
        std::vector<SState> search;
        // 1-based binary heap of (f, state id): the key travels with the entry so that sifting reads one
        // contiguous array instead of one search state per comparison (same comparisons as intrusive_heap.h)
        struct HeapEntry { unsigned int f; int id; };
        std::vector<HeapEntry> open;
        int expanding;         // state popped this round
        bool done;
        QueryResult result;
        // (expanded state, lattice id of its first valid successor that satisfied the goal): what extractPath's
        // "cheapest valid goal action" search finds, every action costing the same (manip_lattice.cpp:2098-2124)
        std::vector<std::pair<int, int>> goal_succ;
        // this round's successors (filled by expandOne, consumed by absorbOne)
        std::vector<double> succ_q1;
        std::vector<int> succ_coord;   // absorbOne scratch: lattice coordinates / table slots of the valid successors
        std::vector<uint64_t> succ_hslot;   // full hash of successor e
        std::vector<int> succ_id;
        int n_succ;
        int edge_begin;
    };

    // minimal fork-join pool: run(f) calls f(tid) on every thread (the caller is tid 0)
    class Pool
    {
    public:
        explicit Pool(int n);
        ~Pool();
        void run(const std::function<void(int)>& f);
        int size() const { return m_n; }
    private:
        int m_n;
        std::vector<std::thread> m_threads;
        std::atomic<int> m_generation{ 0 };
        std::atomic<int> m_done{ 0 };
        std::atomic<bool> m_stop{ false };
        std::atomic<int> m_sleepers{ 0 };
        std::mutex m_mutex;
        std::condition_variable m_cv;
        const std::function<void(int)>* m_job = nullptr;
        void worker(int tid);
    };

    smplgpu_ctx* m_ctx;
    PlannerConfig m_cfg;
    int m_max_concurrent;
    std::vector<std::vector<double>> m_prim_deltas;
    std::vector<bool> m_prim_short;
    std::vector<double> m_coord_deltas;
    std::vector<int> m_coord_vals;
    BatchStats m_stats;

    void stateToCoord(const double* q, std::vector<int>& coord) const;
    bool checkJointLimits(const double* q) const;
    void worldToGrid(const double* p, int* cell) const;
    int computeKey(const SState& s) const;
    SState& sstate(Query& Q, int id);
    void touch(Query& Q, int id);
    void heapPush(Query& Q, int id);
    void heapPop(Query& Q);
    void percolateUp(Query& Q, size_t pivot);
    void percolateDown(Query& Q, size_t pivot);
    void finish(Query& Q, bool found);
    void initQuery(Query& Q, int index, int slot, const double* goal);
    void expandOne(Query& Q);   // pop the next state and generate its in-limit successors (no device work)
    void absorbOne(Query& Q, const uint8_t* verdict, const int32_t* h, const int32_t* gd, const double* off);
};

} // namespace smplhost

#endif
