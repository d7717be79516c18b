// Exact (IEEE double) resolution of the items the certified single-precision pass could not decide, one WARP per
// item.  The thread-per-item kernels of validity.cuh spend ~30 us (a state) / ~60 us (an edge) on a list of a few
// thousand items however short it is: one thread walks 14 links of double-precision sin/cos and 3x4 products, then
// every sphere tree, then every tree pair, all dependent.  Here the independent parts run side by side:
//   1. lane l computes the joint transform J_l of link l (the sin/cos);
//   2. the chain T_l = T_parent * J_l runs link by link, but the 12 entries of a product are 12 lanes;
//   3. lane t descends sphere tree t against the field; 4. lane p descends tree pair p.
// Every number is produced by the same expression, in the same order, as in validity.cuh (joint_transform, the
// dot3-based xf_mul entries, xf_point, df_lookup, the pair descent), so the verdicts are bit-identical to it and to
// the reference; only WHICH lane evaluates an entry differs.  Verdicts are ANDs, so visiting order is free.
#pragma once

#include "model.cuh"
#include "validity.cuh"

namespace smplgpu {

constexpr int RESOLVE_WARPS = 4;   // warps (= items in flight) per block

// One state by one warp.  T: shared memory, MAX_LINKS x 12 doubles owned by this warp.  Returns the verdict (uniform).
__device__ __forceinline__ bool check_state_warp(const DevModel* __restrict__ M, const uint16_t* __restrict__ df,
                                                 const GridParams& G, const double* __restrict__ qa,
                                                 const double* __restrict__ qb, double alpha, double* T, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const int nl = M->n_links;
    // 1. joint transforms (MotionInterpolation::interpolate for the joint value, as check_state)
    for (int l = lane; l < nl; l += 32) {
        double val;
        const int v = M->link_var[l];
        if (v >= 0) {
            const double a = qa[v];
            if (qb != nullptr) {
                const double b = qb[v];
                const double diff = (M->var_type[v] == 1) ? normalize_angle(b - a) : (b - a);
                val = a + alpha * diff;
            } else {
                val = a;
            }
        } else {
            val = M->link_const[l];
        }
        Xf J;
        joint_transform(M, l, val, J);
#pragma unroll
        for (int i = 0; i < 12; ++i) T[l * 12 + i] = J.m[i];
    }
    __syncwarp();
    // 2. T_l = T_parent * J_l, parents first; entry (i, j) of xf_mul by lane 4 i + j
    for (int l = 0; l < nl; ++l) {
        const int p = M->link_parent[l];
        double out = 0.0;
        if (lane < 12) {
            const int i = lane >> 2, j = lane & 3;
            const double* A = p < 0 ? M->link_base[l] : &T[p * 12];
            const double* B = &T[l * 12];
            if (j < 3) {
                out = dot3(A[4 * i], B[j], A[4 * i + 1], B[4 + j], A[4 * i + 2], B[8 + j]);
            } else {
                out = dot3(A[4 * i], B[3], A[4 * i + 1], B[7], A[4 * i + 2], B[11]) + A[4 * i + 3];
            }
        }
        __syncwarp();
        if (lane < 12) {
            T[l * 12 + lane] = out;
        }
        __syncwarp();
    }
    // 3. sphere trees vs the field, a tree per lane
    bool fail = false;
    {
        int stack[MAX_TREE_DEPTH];
        for (int t = lane; t < M->n_trees && !fail; t += 32) {
            int sp = 0;
            stack[sp++] = M->tree_root[t];
            while (sp > 0) {
                const int node = stack[--sp];
                const double* Tl = &T[M->node_link[node] * 12];
                const double cx = M->node_center[node][0], cy = M->node_center[node][1], cz = M->node_center[node][2];
                const double x = dot3(Tl[0], cx, Tl[1], cy, Tl[2], cz) + Tl[3];
                const double y = dot3(Tl[4], cx, Tl[5], cy, Tl[6], cz) + Tl[7];
                const double z = dot3(Tl[8], cx, Tl[9], cy, Tl[10], cz) + Tl[11];
                const int d2 = df_lookup(df, G, x, y, z);
                if (d2 >= M->node_thresh[node]) {
                    continue;
                }
                const int left = M->node_left[node];
                if (left < 0) {
                    fail = true;
                    break;
                }
                stack[sp++] = left;
                stack[sp++] = M->node_right[node];
            }
        }
        if (__any_sync(FULL, fail)) {
            return false;
        }
        // 4. sphere-tree pairs, a pair per lane
        for (int pi = lane; pi < M->n_pairs && !fail; pi += 32) {
            int sp = 0;
            stack[sp++] = (M->tree_root[M->pair_a[pi]] << 16) | M->tree_root[M->pair_b[pi]];
            while (sp > 0) {
                const int packed = stack[--sp];
                const int n1 = packed >> 16, n2 = packed & 0xFFFF;
                const double* T1 = &T[M->node_link[n1] * 12];
                const double* T2 = &T[M->node_link[n2] * 12];
                const double* c1 = M->node_center[n1];
                const double* c2 = M->node_center[n2];
                const double x1 = dot3(T1[0], c1[0], T1[1], c1[1], T1[2], c1[2]) + T1[3];
                const double y1 = dot3(T1[4], c1[0], T1[5], c1[1], T1[6], c1[2]) + T1[7];
                const double z1 = dot3(T1[8], c1[0], T1[9], c1[1], T1[10], c1[2]) + T1[11];
                const double x2 = dot3(T2[0], c2[0], T2[1], c2[1], T2[2], c2[2]) + T2[3];
                const double y2 = dot3(T2[4], c2[0], T2[5], c2[1], T2[6], c2[2]) + T2[7];
                const double z2 = dot3(T2[8], c2[0], T2[9], c2[1], T2[10], c2[2]) + T2[11];
                const double dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
                const double cd2 = (dx * dx + dy * dy) + dz * dz;
                const double r1 = M->node_radius[n1], r2 = M->node_radius[n2];
                const double rr = r1 + r2;
                if (cd2 > rr * rr) {
                    continue;
                }
                const int l1 = M->node_left[n1], l2 = M->node_left[n2];
                if (l1 < 0 && l2 < 0) {
                    bool allowed = false;
                    for (int k = 0; k < M->n_allowed; ++k) {
                        const int a = M->allowed_a[k], b = M->allowed_b[k];
                        allowed |= (a == n1 && b == n2) || (a == n2 && b == n1);
                    }
                    if (!allowed) {
                        fail = true;
                        break;
                    }
                    continue;
                }
                bool split1;
                if (l1 < 0) {
                    split1 = false;
                } else if (l2 < 0) {
                    split1 = true;
                } else {
                    split1 = r1 > r2;
                }
                if (split1) {
                    stack[sp++] = (l1 << 16) | n2;
                    stack[sp++] = (M->node_right[n1] << 16) | n2;
                } else {
                    stack[sp++] = (n1 << 16) | l2;
                    stack[sp++] = (n1 << 16) | M->node_right[n2];
                }
            }
        }
    }
    return !__any_sync(FULL, fail);
}

// the states on `list` (the single-precision pass's undecided ones), a warp each
__global__ void __launch_bounds__(32 * RESOLVE_WARPS)
states_resolve_kernel(const DevModel* __restrict__ M, const uint16_t* __restrict__ df, GridParams G,
                      const double* __restrict__ q, int n, uint8_t* __restrict__ verdict,
                      const int* __restrict__ list, const int* __restrict__ list_n)
{
    __shared__ double sT[RESOLVE_WARPS][MAX_LINKS * 12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = min(*list_n, n);
    const int n_warps = (int)gridDim.x * RESOLVE_WARPS;
    for (int k = (int)blockIdx.x * RESOLVE_WARPS + warp; k < total; k += n_warps) {
        const int i = list[k];
        const bool ok = check_state_warp(M, df, G, q + (size_t)i * M->dof, nullptr, 0.0, sT[warp], lane);
        if (lane == 0) {
            verdict[i] = ok ? 1 : 0;
        }
        __syncwarp();
    }
}

// the edges on `list`, a warp each: the waypoints the single-precision pass could not decide (bit min(w, 31) of
// unc_mask[edge]; every waypoint when there is no mask), in turn until one fails (collision_space.cpp:538-581) -- the
// others were found valid there, or the edge would not be on the list
__global__ void __launch_bounds__(32 * RESOLVE_WARPS)
edges_resolve_kernel(const DevModel* __restrict__ M, const uint16_t* __restrict__ df, GridParams G,
                     const double* __restrict__ q0, const double* __restrict__ q1, int n,
                     uint8_t* __restrict__ verdict, const int* __restrict__ list, const int* __restrict__ list_n,
                     const int* __restrict__ unc_mask)
{
    __shared__ double sT[RESOLVE_WARPS][MAX_LINKS * 12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = min(*list_n, n);
    const int n_warps = (int)gridDim.x * RESOLVE_WARPS;
    const int dof = M->dof;
    for (int k = (int)blockIdx.x * RESOLVE_WARPS + warp; k < total; k += n_warps) {
        const int i = list[k];
        const unsigned mask = unc_mask != nullptr ? (unsigned)unc_mask[i] : 0xFFFFFFFFu;
        const double* a = q0 + (size_t)i * dof;
        const double* b = q1 + (size_t)i * dof;
        // RobotMotionCollisionModel::getMaxSphereMotion + setWaypointCount (every lane, same value)
        double motion = 0.0;
        for (int v = 0; v < dof; ++v) {
            const int ty = M->var_type[v];
            if (ty == 1) {
                motion += M->var_weight[v] * fabs(normalize_angle(b[v] - a[v]));
            } else if (ty == 0) {
                motion += M->var_weight[v] * fabs(b[v] - a[v]);
            } else {
                motion += fabs(b[v] - a[v]);
            }
        }
        int count = 0;
        if (motion != 0.0) {
            count = max(2, (int)ceil(motion / 0.05) + 1);
        }
        bool ok = true;
        const double inv = count > 1 ? 1.0 / (double)(count - 1) : 0.0;
        for (int w = 0; w < count && ok; ++w) {
            if (!((mask >> min(w, 31)) & 1u)) {
                continue;   // decided (valid) in single precision
            }
            ok = check_state_warp(M, df, G, a, b, (double)w * inv, sT[warp], lane);
            __syncwarp();
        }
        if (lane == 0) {
            verdict[i] = ok ? 1 : 0;
        }
    }
}

} // namespace smplgpu
