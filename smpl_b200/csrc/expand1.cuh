// One expansion of ONE lattice state, speculatively: everything the reference's ManipLattice::GetSuccs and
// ARAStar::expand will ask the three plug-ins about a parent state and its motion-primitive successors, computed
// by a single launch and written straight into page-locked host memory.
//
// The reference walks an expansion through ~65 virtual calls:
//   ManipLatticeActionSpace::apply      computePlanningLinkFK(parent), getMetricGoalDistance(link position)
//                                                                     manip_lattice_action_space.cpp:385-396
//   ManipLattice::checkAction, per action   checkJointLimits(successor), isStateToStateValid(parent, successor)
//                                                                     manip_lattice.cpp:1520, 1549
//   ManipLattice::isGoal, per valid action  computePlanningFrameFK(successor)      manip_lattice.cpp:1582-1640
//   ARAStar::reinitSearchState, per new id  GetGoalHeuristic(successor)            arastar.cpp:613-618
// Answered one by one they are ~65 launches and PCIe round trips.  The adapters (smpl_b200/host/gpu_adapters.cpp)
// instead call smplgpu_expand_state on the first question about a state they have no answers for, and serve the
// following calls from the record this kernel leaves behind.  Same answers, callers untouched.
//
// Latency, not throughput, is what matters for one expansion (~25 edges of 2-6 waypoints): one block per successor --
// warp 0 checks the edge parent -> successor, one waypoint per lane, ANDed with a warp vote; lane 0 of warp 1
// does the successor's joint limits, planning-frame FK and BFS lookups meanwhile -- plus one block for the
// parent itself.  A waypoint is checked in certified single precision (validity32.cuh: 7 folded links in float
// instead of 14 in double) and, in the same lane and right away, in double when that cannot decide it: no resolve
// pass, no second launch.  The last block to finish publishes a sequence number the host spins on.
#pragma once

#include "heuristic.cuh"
#include "model.cuh"
#include "validity.cuh"
#include "validity32.cuh"

namespace smplgpu {

constexpr int EXPAND1_THREADS = 64;

struct Expand1Parent { double q[MAX_DOF]; };

static_assert(SMPLGPU_MAX_DOF == MAX_DOF, "smplgpu_succ_info::state and the device tables disagree on the dof bound");

// dynamic shared memory: double-precision slots (n_slots * 12 * blockDim doubles) | with a single-precision model:
// its blob | per-thread f32 slots and root centres
__global__ void __launch_bounds__(EXPAND1_THREADS)
expand_state_kernel(const DevModel* __restrict__ M, const uint16_t* __restrict__ df, GridParams G,
                    const float* __restrict__ blob_g, int blob_words, Grid32 G32,
                    const int* __restrict__ bfs, int dimx, int dimy, int dimz,
                    Expand1Parent parent, const double* __restrict__ deltas, int cost_per_cell,
                    smplgpu_succ_info* out, unsigned int* done, unsigned long long* flag, unsigned long long seq)
{
    extern __shared__ double smem[];
    __shared__ double s_q0[MAX_DOF], s_q1[MAX_DOF];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* blob = reinterpret_cast<float*>(smem + (size_t)M->n_slots * 12 * blockDim.x);
    if (blob_g != nullptr) {
        copy_blob(blob, blob_g, blob_words);
    }
    const int b = blockIdx.x;            // 0 = the parent itself, b >= 1 = motion primitive b - 1
    const int dof = M->dof;
    if (tid < dof) {
        const double a = parent.q[tid];
        s_q0[tid] = a;
        // ManipLatticeActionSpace::applyMotionPrimitive: one IEEE addition per joint
        s_q1[tid] = b > 0 ? deltas[(size_t)(b - 1) * dof + tid] + a : a;
    }
    __syncthreads();
    smplgpu_succ_info* o = out + b;

    if (warp == 1) {
        if (lane < dof) {
            o->state[lane] = s_q1[lane];
        }
        if (lane == 0) {
            o->limits_ok = joint_limits_ok(M, s_q1) ? 1 : 0;
            o->is_parent = b == 0 ? 1 : 0;
            double pose[6], link[3];
            planning_frame_fk(M, s_q1, pose, link);
#pragma unroll
            for (int k = 0; k < 6; ++k) o->pose[k] = pose[k];
            o->link_xyz[0] = link[0]; o->link_xyz[1] = link[1]; o->link_xyz[2] = link[2];
            int h = 0, gd = 0x7FFFFFFF;
            if (bfs != nullptr) {
                bool inb;
                const int d_off = bank_lookup(bfs, dimx, dimy, dimz, 0, G, pose[0], pose[1], pose[2], inb);
                h = (!inb || d_off == 0x7FFFFFFF) ? 32767 : cost_per_cell * d_off;
                gd = bank_lookup(bfs, dimx, dimy, dimz, 0, G, link[0], link[1], link[2], inb);
                if (!inb) {
                    gd = -2;   // SMPLGPU_BFS_OUT_OF_BOUNDS, as smplgpu_bfs_distances reports it
                }
            }
            o->h = h;
            o->goal_dist_cells = gd;
        }
    } else {
        Counters cnt = { 0u, 0u, 0u };
        S32 S;
        float* slots32 = blob + blob_words;
        if (blob_g != nullptr) {
            S = view32(blob);
        }
        // one state: certified single precision first, double in the same lane when that cannot decide
        auto state_ok = [&](const double* qa, const double* qb, double alpha) {
            if (blob_g != nullptr) {
                const int r = check_state32(S, M->var_type, df, G32, qa, qb, alpha, slots32, cnt);
                if (r != 2) {
                    return r == 1;
                }
            }
            return check_state(M, df, G, qa, qb, alpha, smem, cnt);
        };
        if (b == 0) {
            // CollisionSpace::isStateValid(parent)
            bool ok = true;
            if (lane == 0) {
                ok = state_ok(s_q1, nullptr, 0.0);
            }
            ok = __all_sync(0xffffffffu, ok);
            if (lane == 0) {
                o->state_valid = ok ? 1 : 0;
                o->edge_valid = ok ? 1 : 0;
                o->waypoints = 0;
            }
        } else {
            // CollisionSpace::isStateToStateValid(parent, successor): getMaxSphereMotion + setWaypointCount
            double motion = 0.0;
            for (int v = 0; v < dof; ++v) {
                const int ty = M->var_type[v];
                if (ty == 1) {
                    motion += M->var_weight[v] * fabs(normalize_angle(s_q1[v] - s_q0[v]));
                } else if (ty == 0) {
                    motion += M->var_weight[v] * fabs(s_q1[v] - s_q0[v]);
                } else {
                    motion += fabs(s_q1[v] - s_q0[v]);
                }
            }
            int count = 0;
            if (motion != 0.0) {
                count = max(2, (int)ceil(motion / 0.05) + 1);
            }
            bool ok = true;
            if (count > 0) {
                const double inv = 1.0 / (double)(count - 1);
                for (int w0 = 0; w0 < count && ok; w0 += 32) {
                    const int w = w0 + lane;
                    bool mine = true;
                    if (w < count) {
                        mine = state_ok(s_q0, s_q1, (double)w * inv);
                    }
                    ok = __all_sync(0xffffffffu, mine);
                }
            }
            if (lane == 0) {
                o->state_valid = 0;   // not asked for successors (the edge's last waypoint covers it)
                o->edge_valid = ok ? 1 : 0;
                o->waypoints = count;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            *((volatile unsigned long long*)flag) = seq;
        }
    }
}

} // namespace smplgpu
