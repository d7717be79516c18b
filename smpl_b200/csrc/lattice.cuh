// The manipulation lattice of MANY concurrent queries, resident on the device: successor generation, joint limits,
// stateToCoord, the coordinate hash table (getOrCreateState), the goal test and the heuristic of one expansion round
// run here, so that a round ships 8 bytes per expansion (bank slot, state id) and gets back one int pair per
// successor (state id | goal flag, heuristic).  The host keeps what north_star says it keeps: ARA*'s OPEN list and
// search states, indexed by the ids this table hands out.
//
// Reference semantics restated (file:line under dyouakim/smpl):
//   ManipLattice::GetSuccs                          smpl/src/graph/manip_lattice.cpp:219-313
//   ManipLatticeActionSpace::apply / mprimActive    smpl/src/graph/manip_lattice_action_space.cpp:376-449, 662-691
//   ManipLattice::checkAction (joint limits first)  manip_lattice.cpp:1511-1580
//   ManipLattice::stateToCoord                      manip_lattice.cpp:1263-1289
//   ManipLattice::getOrCreateState / createHashEntry  manip_lattice.cpp:1291-1356  (ids in creation order)
//   ManipLattice::isGoal, XYZ_GOAL                  manip_lattice.cpp:1673-1687
//   BfsHeuristic::GetGoalHeuristic                  smpl/src/heuristic/bfs_heuristic.cpp:148-163
// State ids must come out in the reference's creation order (they are what plans are compared by): an expansion's
// successors are entered one after the other in primitive order by ONE warp, and a query has at most one
// expansion per round, so no two warps ever touch the same query's table.
#pragma once

#include "heuristic.cuh"
#include "model.cuh"
#include "validity.cuh"

namespace smplgpu {

constexpr int LATTICE_MAX_STRIDE = 32;        // successors per expansion: one lane each
constexpr int LATTICE_GOAL_FLAG = 1 << 30;    // in a successor word: the action reaches the goal region

struct LatticeBank
{
    int n_slots, cap, table_size;             // states per slot, hash slots per slot (power of two)
    int stride;                               // successor words per expansion = max(#long, #short primitives)
    int n_long, n_short;
    int use_short_dist;
    double short_dist_thresh, res;
    double tol[3];
    int cost_per_cell;
    double* q;        // [n_slots][cap][dof]
    int* coord;       // [n_slots][cap][dof]
    int* gdist;       // [n_slots][cap]   BFS cells at the planning link (getMetricGoalDistance / res)
    int* table;       // [n_slots][table_size] state id, -1 = empty
    int* count;       // [n_slots] lattice size (ids 0 .. count-1; 0 = the goal state, 1 = the start state)
    double* goal;     // [n_slots][3]
    const double* deltas;     // [n_prims][dof]
    const int* long_list;     // primitive indices in table order
    const int* short_list;
};

// ManipLattice::stateToCoord (manip_lattice.cpp:1263-1289; every KDL planning variable is continuous or bounded)
__device__ __forceinline__ int state_to_coord(const DevModel* __restrict__ M, const LatticeParams& L, const int* vals,
                                              int v, double q)
{
    const double PI = 3.14159265358979323846;
    if (M->var_type[v] == 1) {
        double pos = normalize_angle(q);
        if (pos < 0.0) {
            pos += 2.0 * PI;
        }
        int c = __double2int_rz((pos + L.delta[v] * 0.5) / L.delta[v]);
        if (c == vals[v]) {
            c = 0;
        }
        return c;
    }
    return __double2int_rz(((q - M->var_min[v]) / L.delta[v]) + 0.5);
}

struct LatticeVals { int v[MAX_DOF]; };

__device__ __forceinline__ unsigned int coord_hash(const int* c, int dof)
{
    unsigned long long x = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < dof; ++i) {
        x ^= (unsigned long long)(unsigned int)c[i] + 0x9E3779B97F4A7C15ull + (x << 6) + (x >> 2);
    }
    return (unsigned int)(x ^ (x >> 32));
}

// tables of the listed slots -> empty
__global__ void lattice_clear_kernel(LatticeBank B, const int* __restrict__ slots, int n)
{
    const size_t total = (size_t)n * B.table_size;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / B.table_size);
        B.table[(size_t)slots[k] * B.table_size + (i - (size_t)k * B.table_size)] = -1;
    }
}

// ManipLattice::setGoal + setStart bookkeeping: id 0 = the goal state (no coordinates), id 1 = the start state
__global__ void lattice_begin_kernel(const DevModel* __restrict__ M, LatticeBank B, LatticeParams L, LatticeVals V,
                                     const int* __restrict__ slots, const double* __restrict__ starts,
                                     const int* __restrict__ start_gd, const double* __restrict__ goals, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) {
        return;
    }
    const int s = slots[k];
    const int dof = M->dof;
    int c[MAX_DOF];
    double* q1 = B.q + ((size_t)s * B.cap + 1) * dof;
    int* c1 = B.coord + ((size_t)s * B.cap + 1) * dof;
    for (int v = 0; v < dof; ++v) {
        const double a = starts[(size_t)k * dof + v];
        q1[v] = a;
        c[v] = state_to_coord(M, L, V.v, v, a);
        c1[v] = c[v];
        B.q[((size_t)s * B.cap) * dof + v] = 0.0;
    }
    B.gdist[(size_t)s * B.cap] = 0;
    B.gdist[(size_t)s * B.cap + 1] = start_gd[k];
    B.table[(size_t)s * B.table_size + (coord_hash(c, dof) & (unsigned int)(B.table_size - 1))] = 1;
    B.count[s] = 2;
    B.goal[3 * s] = goals[3 * k];
    B.goal[3 * s + 1] = goals[3 * k + 1];
    B.goal[3 * s + 2] = goals[3 * k + 2];
}

// Round, step 1: the edges of every expansion.  Thread (i, j): j-th active primitive of the state expansion i pops.
// Inactive positions and successors beyond a joint limit become zero-length edges (no waypoint, no work in the
// edge kernels) and are flagged.
__global__ void lattice_gen_kernel(const DevModel* __restrict__ M, LatticeBank B, const int* __restrict__ slot,
                                   const int* __restrict__ parent, int n, double* __restrict__ q0,
                                   double* __restrict__ q1, uint8_t* __restrict__ active, unsigned long long* stats)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 4) {
        stats[t] = 0;   // counters of the edge kernels that follow (saves a memset call per round)
    }
    if (t >= n * B.stride) {
        return;
    }
    const int i = t / B.stride, j = t - i * B.stride;
    const int s = slot[i], p = parent[i];
    const int dof = M->dof;
    const double* a = B.q + ((size_t)s * B.cap + p) * dof;
    // mprimActive: short-distance primitives once the planning link is within the threshold of the goal
    const double goal_dist = (double)B.gdist[(size_t)s * B.cap + p] * B.res;
    const bool near_goal = B.use_short_dist && goal_dist <= B.short_dist_thresh;
    const int cnt = near_goal ? B.n_short : B.n_long;
    double succ[MAX_DOF];
    bool on = j < cnt;
    if (on) {
        const int prim = near_goal ? B.short_list[j] : B.long_list[j];
        for (int v = 0; v < dof; ++v) {
            succ[v] = B.deltas[(size_t)prim * dof + v] + a[v];
        }
        on = joint_limits_ok(M, succ);
    }
    for (int v = 0; v < dof; ++v) {
        const double av = a[v];
        q0[(size_t)t * dof + v] = av;
        q1[(size_t)t * dof + v] = on ? succ[v] : av;
    }
    active[t] = on ? 1 : 0;
}

// Round, step 3 (after the edge kernels): one warp per expansion, lane j = successor j.  Valid successors get their
// planning-frame pose (goal test, heuristic, metric goal distance), their coordinates and hash in parallel; then
// they are looked up / entered ONE AFTER THE OTHER in primitive order, so ids come out in the reference's
// creation order.  out_succ[i][j] = id | LATTICE_GOAL_FLAG, or -1; out_h[i][j] = GetGoalHeuristic(successor);
// out_count[i] = lattice size of the query afterwards.
__global__ void __launch_bounds__(128)
lattice_commit_kernel(const DevModel* __restrict__ M, GridParams G, LatticeBank B, LatticeParams L, LatticeVals V,
                      const int* __restrict__ bfs, int dimx, int dimy, int slot_dimz,
                      const int* __restrict__ slot, int n, const double* __restrict__ q1,
                      const uint8_t* __restrict__ active, const uint8_t* __restrict__ verdict,
                      int* __restrict__ out_succ, int* __restrict__ out_h, int* __restrict__ out_count,
                      const unsigned long long* __restrict__ stats)
{
    const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) {
        return;
    }
    const int s = slot[i];
    const int dof = M->dof;
    const int t = i * B.stride + lane;
    const bool valid = lane < B.stride && active[t] != 0 && verdict[t] != 0;
    int c[MAX_DOF];
    unsigned int hv = 0;
    int h = 0, gd = 0;
    bool is_goal = false;
    const double* q = q1 + (size_t)t * dof;
    if (valid) {
        double pose[6], link[3];
        planning_frame_fk(M, q, pose, link);
        bool inb;
        const int d_off = bank_lookup(bfs, dimx, dimy, slot_dimz, s, G, pose[0], pose[1], pose[2], inb);
        h = (!inb || d_off == 0x7FFFFFFF) ? 32767 : B.cost_per_cell * d_off;
        gd = bank_lookup(bfs, dimx, dimy, slot_dimz, s, G, link[0], link[1], link[2], inb);
        is_goal = fabs(pose[0] - B.goal[3 * s]) <= B.tol[0] && fabs(pose[1] - B.goal[3 * s + 1]) <= B.tol[1] &&
                  fabs(pose[2] - B.goal[3 * s + 2]) <= B.tol[2];
        for (int v = 0; v < dof; ++v) {
            c[v] = state_to_coord(M, L, V.v, v, q[v]);
        }
        hv = coord_hash(c, dof);
    }
    int id = -1;
    bool full = false;
    unsigned int todo = __ballot_sync(0xffffffffu, valid);
    int* table = B.table + (size_t)s * B.table_size;
    const unsigned int mask = (unsigned int)(B.table_size - 1);
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        if (lane == j) {
            unsigned int slot_i = hv & mask;
            for (;;) {
                const int cand = table[slot_i];
                if (cand < 0) {
                    break;
                }
                const int* cc = B.coord + ((size_t)s * B.cap + cand) * dof;
                bool same = true;
                for (int v = 0; v < dof; ++v) {
                    same = same && cc[v] == c[v];
                }
                if (same) {
                    id = cand;
                    break;
                }
                slot_i = (slot_i + 1) & mask;
            }
            if (id < 0) {
                const int fresh = B.count[s];
                if (fresh >= B.cap) {
                    full = true;
                } else {
                    B.count[s] = fresh + 1;
                    double* qq = B.q + ((size_t)s * B.cap + fresh) * dof;
                    int* cc = B.coord + ((size_t)s * B.cap + fresh) * dof;
                    for (int v = 0; v < dof; ++v) {
                        qq[v] = q[v];
                        cc[v] = c[v];
                    }
                    B.gdist[(size_t)s * B.cap + fresh] = gd;
                    table[slot_i] = fresh;
                    id = fresh;
                }
            }
        }
        __syncwarp();
    }
    if (lane < B.stride) {
        out_succ[t] = id < 0 ? -1 : (id | (is_goal ? LATTICE_GOAL_FLAG : 0));
        out_h[t] = h;
    }
    full = __any_sync(0xffffffffu, full);
    if (lane == 0) {
        out_count[i] = full ? -1 : B.count[s];   // -1: the query ran out of room (the caller sized it too small)
    }
    if (i == 0 && lane == 0) {
        // edges of this round that the double-precision pass resolved, behind the counts (8-byte slot)
        unsigned long long* tail = reinterpret_cast<unsigned long long*>(out_count + ((n + 1) & ~1));
        *tail = stats[3];
    }
}

// joint values of listed states (path extraction): out[k] = q[slot[k]][id[k]]
__global__ void lattice_gather_kernel(LatticeBank B, int dof, const int* __restrict__ slot, const int* __restrict__ id,
                                      int n, double* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * dof) {
        return;
    }
    const int k = t / dof, v = t - k * dof;
    out[t] = B.q[((size_t)slot[k] * B.cap + id[k]) * dof + v];
}

} // namespace smplgpu
