// Kernels (1)+(2)+(3): batched forward kinematics -> sphere tree vs distance
// field -> self-collision sphere pairs, fused; one thread per state (or edge),
// IEEE double throughout, compiled with --fmad=false so every product and sum
// rounds exactly like the reference's x86-64 build (no FMA contraction).
//
// Reference semantics restated (file:line under dyouakim/smpl):
//   joint transforms      sbpl_collision_checking/src/transform_functions.h:95-258
//   link transform        include/sbpl_collision_checking/robot_collision_state.h:385-431
//   sphere position       robot_collision_state.h:560-581
//   sphere vs field       src/collision_operations.h:68-77, 105-165
//   field lookup          smpl/include/smpl/distance_map/detail/distance_map.hpp:281-300, 520-536
//   sphere pairs + ACM    src/self_collision_model.cpp:1093-1218
//   edge interpolation    src/collision_space.cpp:538-581,
//                         include/.../robot_motion_collision_model.h:173-181, 224-249, 297-321, 352-366
//                         src/robot_motion_collision_model.cpp:371-407; smpl/angles.h:45-99
#pragma once

#include "model.cuh"

namespace smplgpu {

constexpr int VALIDITY_THREADS = 128;

struct Xf { double m[12]; }; // row-major 3x4

__device__ __forceinline__ double dot3(double a0, double b0, double a1, double b1, double a2, double b2)
{
    return (a0 * b0 + a1 * b1) + a2 * b2;
}

// Eigen Affine3d * Affine3d: linear = L1*L2, translation = L1*t2 + t1
__device__ __forceinline__ void xf_mul(const Xf& a, const Xf& b, Xf& r)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            r.m[4 * i + j] = dot3(a.m[4 * i], b.m[j], a.m[4 * i + 1], b.m[4 + j], a.m[4 * i + 2], b.m[8 + j]);
        }
        r.m[4 * i + 3] = dot3(a.m[4 * i], b.m[3], a.m[4 * i + 1], b.m[7], a.m[4 * i + 2], b.m[11]) + a.m[4 * i + 3];
    }
}

__device__ __forceinline__ void xf_point(const Xf& a, double x, double y, double z, double& px, double& py, double& pz)
{
    px = dot3(a.m[0], x, a.m[1], y, a.m[2], z) + a.m[3];
    py = dot3(a.m[4], x, a.m[5], y, a.m[6], z) + a.m[7];
    pz = dot3(a.m[8], x, a.m[9], y, a.m[10], z) + a.m[11];
}

// smpl/angles.h:45-62
__device__ __forceinline__ double normalize_angle(double angle)
{
    const double PI = 3.14159265358979323846;
    if (fabs(angle) > 2.0 * PI) {
        angle = fmod(angle, 2.0 * PI);
    }
    if (angle < -PI) {
        angle += 2.0 * PI;
    }
    if (angle > PI) {
        angle -= 2.0 * PI;
    }
    return angle;
}

// Eigen::AngleAxisd(angle, axis).toRotationMatrix() as a 3x4 with zero translation
__device__ __forceinline__ void angle_axis(double angle, double ax, double ay, double az, Xf& r)
{
    double s, c;
    sincos(angle, &s, &c);
    const double sx = s * ax, sy = s * ay, sz = s * az;
    const double k = 1.0 - c;
    const double cx = k * ax, cy = k * ay, cz = k * az;
    double tmp;
    tmp = cx * ay;
    r.m[1] = tmp - sz;
    r.m[4] = tmp + sz;
    tmp = cx * az;
    r.m[2] = tmp + sy;
    r.m[8] = tmp - sy;
    tmp = cy * az;
    r.m[6] = tmp - sx;
    r.m[9] = tmp + sx;
    r.m[0] = cx * ax + c;
    r.m[5] = cy * ay + c;
    r.m[10] = cz * az + c;
    r.m[3] = 0.0;
    r.m[7] = 0.0;
    r.m[11] = 0.0;
}

// transform_functions.h:95-258
__device__ __forceinline__ void joint_transform(const DevModel* __restrict__ M, int l, double val, Xf& t)
{
    const double* o = M->link_origin[l];
    const int fn = M->link_joint[l];
    if (fn == 0) { // fixed
#pragma unroll
        for (int i = 0; i < 12; ++i) t.m[i] = o[i];
    } else if (fn <= 3) {
        double sth, cth;
        sincos(val, &sth, &cth);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double o0 = o[4 * r], o1 = o[4 * r + 1], o2 = o[4 * r + 2];
            double t0, t1, t2;
            if (fn == 1) {          // X
                t0 = o0;
                t1 = cth * o1 + sth * o2;
                t2 = cth * o2 - sth * o1;
            } else if (fn == 2) {   // Y
                t0 = cth * o0 - sth * o2;
                t1 = o1;
                t2 = sth * o0 + cth * o2;
            } else {                // Z
                t0 = o0 * cth + o1 * sth;
                t1 = o1 * cth - o0 * sth;
                t2 = o2;
            }
            t.m[4 * r] = t0;
            t.m[4 * r + 1] = t1;
            t.m[4 * r + 2] = t2;
            t.m[4 * r + 3] = o[4 * r + 3];
        }
    } else if (fn == 4) { // origin * AngleAxis(q, axis)
        Xf a, oo;
        angle_axis(val, M->link_axis[l][0], M->link_axis[l][1], M->link_axis[l][2], a);
#pragma unroll
        for (int i = 0; i < 12; ++i) oo.m[i] = o[i];
        xf_mul(oo, a, t);
    } else { // prismatic: origin * Translate(0, 0, q)
        Xf a, oo;
#pragma unroll
        for (int i = 0; i < 12; ++i) { oo.m[i] = o[i]; a.m[i] = 0.0; }
        a.m[0] = 1.0; a.m[5] = 1.0; a.m[10] = 1.0; a.m[11] = val;
        xf_mul(oo, a, t);
    }
}

// DistanceMap::worldToGrid + isCellValid + cell d2 (0 when outside)
__device__ __forceinline__ int df_lookup(const uint16_t* __restrict__ df, const GridParams& G, double x, double y, double z)
{
    const int gx = __double2int_rz(G.inv_res * (x - G.ox) + 0.5) - 1;
    const int gy = __double2int_rz(G.inv_res * (y - G.oy) + 0.5) - 1;
    const int gz = __double2int_rz(G.inv_res * (z - G.oz) + 0.5) - 1;
    if ((unsigned)gx >= (unsigned)G.nx || (unsigned)gy >= (unsigned)G.ny || (unsigned)gz >= (unsigned)G.nz) {
        return 0;
    }
    return (int)__ldg(&df[((size_t)gx * G.ny + gy) * G.nz + gz]);
}

struct Counters { unsigned int lookups, pairs, waypoints; };

// per-thread slot storage in shared memory: element e of slot s of thread t
// lives at smem[(s * 12 + e) * blockDim + t]  (conflict-free for a warp)
// (stride, idx) = (threads sharing the slot storage, this thread's column); 0 = the whole block, threadIdx.x
__device__ __forceinline__ void slot_store(double* smem, int slot, const Xf& t, int stride = 0, int idx = 0)
{
    const int st = stride > 0 ? stride : (int)blockDim.x;
    double* p = smem + (size_t)slot * 12 * st + (stride > 0 ? idx : (int)threadIdx.x);
#pragma unroll
    for (int e = 0; e < 12; ++e) p[e * st] = t.m[e];
}

__device__ __forceinline__ void slot_load(const double* smem, int slot, Xf& t, int stride = 0, int idx = 0)
{
    const int st = stride > 0 ? stride : (int)blockDim.x;
    const double* p = smem + (size_t)slot * 12 * st + (stride > 0 ? idx : (int)threadIdx.x);
#pragma unroll
    for (int e = 0; e < 12; ++e) t.m[e] = p[e * st];
}

__device__ __forceinline__ void node_pos(const DevModel* __restrict__ M, const double* smem, int node,
                                         double& x, double& y, double& z, int stride = 0, int idx = 0)
{
    Xf t;
    slot_load(smem, M->link_slot[M->node_link[node]], t, stride, idx);
    xf_point(t, M->node_center[node][0], M->node_center[node][1], M->node_center[node][2], x, y, z);
}

// One state: FK, sphere trees vs field (inline, while T_link is in registers),
// then tree pairs.  Returns true when the state is valid.
//   qa/qb/alpha: state value of planning variable v is
//       qb == nullptr ? qa[v] : qa[v] + alpha * diff(v)     (MotionInterpolation::interpolate)
__device__ bool check_state(const DevModel* __restrict__ M, const uint16_t* __restrict__ df, const GridParams& G,
                            const double* __restrict__ qa, const double* __restrict__ qb, double alpha,
                            double* smem, Counters& cnt, int sstride = 0, int sidx = 0)
{
    Xf T;  // transform of the previously processed link
    int stack[MAX_TREE_DEPTH];

    const int nl = M->n_links;
    for (int l = 0; l < nl; ++l) {
        // joint value
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

        Xf J, P;
        joint_transform(M, l, val, J);
        const int p = M->link_parent[l];
        if (p < 0) {
#pragma unroll
            for (int i = 0; i < 12; ++i) P.m[i] = M->link_base[l][i];
        } else if (p == l - 1) {
            P = T;
        } else {
            slot_load(smem, M->link_slot[p], P, sstride, sidx);
        }
        xf_mul(P, J, T);
        const int slot = M->link_slot[l];
        if (slot >= 0) {
            slot_store(smem, slot, T, sstride, sidx);
        }

        // sphere trees rooted on this link vs the distance field
        for (int ti = M->link_tree_begin[l]; ti < M->link_tree_end[l]; ++ti) {
            int sp = 0;
            stack[sp++] = M->tree_root[M->tree_by_link[ti]];
            while (sp > 0) {
                const int node = stack[--sp];
                double x, y, z;
                xf_point(T, M->node_center[node][0], M->node_center[node][1], M->node_center[node][2], x, y, z);
                ++cnt.lookups;
                const int d2 = df_lookup(df, G, x, y, z);
                if (d2 >= M->node_thresh[node]) {
                    continue;
                }
                const int left = M->node_left[node];
                if (left < 0) {
                    return false; // failing leaf
                }
                stack[sp++] = left;
                stack[sp++] = M->node_right[node];
            }
        }
    }

    // sphere-tree pairs
    const int np = M->n_pairs;
    for (int pi = 0; pi < np; ++pi) {
        int sp = 0;
        stack[sp++] = (M->tree_root[M->pair_a[pi]] << 16) | M->tree_root[M->pair_b[pi]];
        while (sp > 0) {
            const int packed = stack[--sp];
            const int n1 = packed >> 16, n2 = packed & 0xFFFF;
            double x1, y1, z1, x2, y2, z2;
            node_pos(M, smem, n1, x1, y1, z1, sstride, sidx);
            node_pos(M, smem, n2, x2, y2, z2, sstride, sidx);
            ++cnt.pairs;
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
                    return false;
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
    return true;
}

__device__ __forceinline__ void flush_counters(const Counters& c, unsigned long long* stats)
{
    if (stats == nullptr) {
        return;
    }
    unsigned int a = c.lookups, b = c.pairs, w = c.waypoints;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        w += __shfl_xor_sync(0xffffffffu, w, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&stats[0], (unsigned long long)a);
        atomicAdd(&stats[1], (unsigned long long)b);
        atomicAdd(&stats[2], (unsigned long long)w);
    }
}

// CollisionSpace::isStateValid, batched.  With `list` the kernel resolves only the states the
// single-precision pass (validity32.cuh) could not decide: item k is state list[k], k < *list_n.
__global__ void __launch_bounds__(VALIDITY_THREADS)
states_valid_kernel(const DevModel* __restrict__ M, const uint16_t* __restrict__ df, GridParams G,
                    const double* __restrict__ q, int n, uint8_t* __restrict__ verdict,
                    unsigned long long* stats, const int* __restrict__ list, const int* __restrict__ list_n)
{
    extern __shared__ double smem[];
    const int total = list != nullptr ? min(*list_n, n) : n;
    Counters cnt = { 0u, 0u, 0u };
    // With a list (the few states the single-precision pass could not decide) the items are dealt one per WARP
    // first: these states take different branches all along the chain and the descent, so 32 of them in one warp
    // would run one after the other; spread out, a warp holds two or three.
    const int n_warps = (int)(gridDim.x * (blockDim.x >> 5));
    const int k0 = list != nullptr ? (int)(threadIdx.x & 31) * n_warps + (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5))
                                   : (int)(blockIdx.x * blockDim.x + threadIdx.x);
    for (int k = k0; k < total; k += gridDim.x * blockDim.x) {
        const int i = list != nullptr ? list[k] : k;
        ++cnt.waypoints;
        const bool ok = check_state(M, df, G, q + (size_t)i * M->dof, nullptr, 0.0, smem, cnt);
        verdict[i] = ok ? 1 : 0;
    }
    flush_counters(cnt, stats);
}

// CollisionSpace::isStateToStateValid, batched.  Waypoints are generated on the
// device and spread over the block: each thread computes the waypoint count of
// its own edge, a block-wide exclusive scan turns the counts into offsets, and
// the block's threads then walk the flattened (edge, waypoint) list, so lanes
// stay busy whatever the mix of short and long edges.  Consecutive lanes hold
// consecutive waypoints of one edge (similar poses => similar control flow).
// The verdict is the AND over all waypoints (order only affects early-out in
// the reference); a waypoint is skipped once its edge is known to be invalid.
//
// Dynamic shared memory: slot storage (n_slots*12*blockDim doubles) followed by
// (blockDim + 1) ints of offsets, blockDim ints of per-edge verdicts and blockDim edge ids.
__global__ void __launch_bounds__(VALIDITY_THREADS)
edges_valid_kernel(const DevModel* __restrict__ M, const uint16_t* __restrict__ df, GridParams G,
                   const double* __restrict__ q0, const double* __restrict__ q1, int n,
                   uint8_t* __restrict__ verdict, int* __restrict__ counts, unsigned long long* stats,
                   const int* __restrict__ list, const int* __restrict__ list_n)
{
    extern __shared__ double smem[];
    int* s_off = reinterpret_cast<int*>(smem + (size_t)M->n_slots * 12 * blockDim.x);
    int* s_ok = s_off + blockDim.x + 1;
    int* s_eid = s_ok + blockDim.x;
    const int tid = threadIdx.x;
    const int dof = M->dof;
    const int total_edges = list != nullptr ? min(*list_n, n) : n;
    Counters cnt = { 0u, 0u, 0u };

    // with `list` (the few edges the single-precision pass could not decide) a block takes only a quarter
    // as many edges as it has threads, so a thread ends up with about one waypoint and the pass stays short;
    // a block may then walk several chunks
    const int per_block = list != nullptr ? max(1, (int)blockDim.x / 4) : (int)blockDim.x;
    for (int first = blockIdx.x * per_block; first < total_edges; first += gridDim.x * per_block) {
        const int k = first + tid;
        const int i = (tid < per_block && k < total_edges) ? (list != nullptr ? list[k] : k) : -1;
        int count = 0;
        if (i >= 0) {
            const double* a = q0 + (size_t)i * dof;
            const double* b = q1 + (size_t)i * dof;
            // RobotMotionCollisionModel::getMaxSphereMotion(start, finish, variables)
            double motion = 0.0;
            for (int v = 0; v < dof; ++v) {
                const int ty = M->var_type[v];
                double dist;
                if (ty == 1) {          // continuous
                    dist = fabs(normalize_angle(b[v] - a[v]));
                    motion += M->var_weight[v] * dist;
                } else if (ty == 0) {   // revolute
                    dist = fabs(b[v] - a[v]);
                    motion += M->var_weight[v] * dist;
                } else {                // prismatic
                    dist = fabs(b[v] - a[v]);
                    motion += dist;
                }
            }
            // fillMotionInterpolation + setWaypointCount
            if (motion != 0.0) {
                count = max(2, (int)ceil(motion / 0.05) + 1);
            }
            if (counts != nullptr) {
                counts[i] = count;
            }
        }
        // block-wide inclusive scan of the counts (Hillis-Steele over <= 128 entries)
        s_off[tid + 1] = count;
        s_ok[tid] = 1;
        s_eid[tid] = i;
        if (tid == 0) {
            s_off[0] = 0;
        }
        __syncthreads();
        for (int d = 1; d < (int)blockDim.x; d <<= 1) {
            int add = 0;
            if (tid + 1 > d) {
                add = s_off[tid + 1 - d];
            }
            __syncthreads();
            s_off[tid + 1] += add;
            __syncthreads();
        }
        const int total = s_off[blockDim.x];

        for (int item = tid; item < total; item += blockDim.x) {
            // edge of this item: largest e with s_off[e] <= item
            int lo = 0, hi = blockDim.x;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_off[mid] <= item) {
                    lo = mid;
                } else {
                    hi = mid;
                }
            }
            const int e = lo;
            if (*((volatile int*)&s_ok[e]) == 0) {
                continue; // edge already invalid (a stale 1 only costs one redundant waypoint check)
            }
            const int w = item - s_off[e];
            const int cnt_e = s_off[e + 1] - s_off[e];
            const double inv = 1.0 / (double)(cnt_e - 1);     // m_waypoint_count_inv
            const double alpha = (double)w * inv;             // interpolate(n): alpha = n * inv
            const double* a = q0 + (size_t)s_eid[e] * dof;
            const double* b = q1 + (size_t)s_eid[e] * dof;
            ++cnt.waypoints;
            if (!check_state(M, df, G, a, b, alpha, smem, cnt)) {
                atomicAnd(&s_ok[e], 0);   // same-value writes from several lanes: atomic, so racecheck-clean
            }
        }
        __syncthreads();
        if (i >= 0) {
            verdict[i] = s_ok[tid] ? 1 : 0;
        }
        __syncthreads();
    }
    flush_counters(cnt, stats);
}

// ManipLatticeActionSpace::applyMotionPrimitive (manip_lattice_action_space.cpp:575-610) for a batch:
// successor = parent + delta of the edge's motion primitive (one IEEE addition per joint)
__global__ void apply_mprims_kernel(const double* __restrict__ q0, const int* __restrict__ prim,
                                    const double* __restrict__ deltas, int n_prims, int dof, int n,
                                    double* __restrict__ q1)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * dof) {
        return;
    }
    const int e = (int)(i / dof), v = (int)(i - (size_t)e * dof);
    const int p = prim[e];
    const double a = q0[i];
    q1[i] = (p >= 0 && p < n_prims) ? deltas[(size_t)p * dof + v] + a : a;
}

// ManipLattice::coordToState (manip_lattice.cpp:1245-1261) for a batch of lattice coordinates, and -- for edges --
// the successor parent + delta of the edge's motion primitive: a lattice state crosses the bus as dof 16-bit
// coordinates (plus one byte of primitive id per edge) instead of dof doubles.
struct LatticeParams
{
    double base[MAX_DOF];    // m_min_limits for bounded variables, 0 for continuous ones
    double delta[MAX_DOF];   // m_coord_deltas
    int bounded[MAX_DOF];    // adds `base` (the reference's two branches differ in the addition only)
};

__global__ void lattice_states_kernel(const int16_t* __restrict__ coords, const uint8_t* __restrict__ prim,
                                      const double* __restrict__ deltas, int n_prims, LatticeParams L, int dof, int n,
                                      double* __restrict__ q0, double* __restrict__ q1)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * dof) {
        return;
    }
    const int e = (int)(i / dof), v = (int)(i - (size_t)e * dof);
    const int c = coords[i];
    // state[i] = coord[i] * m_coord_deltas[i]   |   m_min_limits[i] + coord[i] * m_coord_deltas[i]
    const double prod = (double)c * L.delta[v];
    const double a = L.bounded[v] ? L.base[v] + prod : prod;
    q0[i] = a;
    if (q1 != nullptr) {
        const int p = prim[e];
        q1[i] = p < n_prims ? deltas[(size_t)p * dof + v] + a : a;
    }
}

// Edges between rows of a point table (path post-processing: every (i, j) pair of a path's points is a candidate
// shortcut, post_processing.cpp:99-121): q0[e] = points[a[e]], q1[e] = points[b[e]]
__global__ void gather_edges_kernel(const double* __restrict__ points, const int* __restrict__ a,
                                    const int* __restrict__ b, int dof, int n,
                                    double* __restrict__ q0, double* __restrict__ q1)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * dof) {
        return;
    }
    const int e = (int)(i / dof), v = (int)(i - (size_t)e * dof);
    q0[i] = points[(size_t)a[e] * dof + v];
    q1[i] = points[(size_t)b[e] * dof + v];
}

// Kernel (1) alone: sphere centres of every tree node, out[n][n_nodes][3]
__global__ void __launch_bounds__(VALIDITY_THREADS)
fk_centers_kernel(const DevModel* __restrict__ M, const double* __restrict__ q, int n, double* __restrict__ out)
{
    extern __shared__ double smem[];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const double* qa = q + (size_t)i * M->dof;
    Xf T;
    for (int l = 0; l < M->n_links; ++l) {
        const int v = M->link_var[l];
        const double val = v >= 0 ? qa[v] : M->link_const[l];
        Xf J, P;
        joint_transform(M, l, val, J);
        const int p = M->link_parent[l];
        if (p < 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) P.m[k] = M->link_base[l][k];
        } else if (p == l - 1) {
            P = T;
        } else {
            slot_load(smem, M->link_slot[p], P);
        }
        xf_mul(P, J, T);
        if (M->link_slot[l] >= 0) {
            slot_store(smem, M->link_slot[l], T);
        }
        // every node on this link
        for (int node = 0; node < M->n_nodes; ++node) {
            if (M->node_link[node] == l) {
                double x, y, z;
                xf_point(T, M->node_center[node][0], M->node_center[node][1], M->node_center[node][2], x, y, z);
                double* o = out + ((size_t)i * M->n_nodes + node) * 3;
                o[0] = x; o[1] = y; o[2] = z;
            }
        }
    }
}

// CollisionSpace::collisionDistance (collision_space.cpp:496-500 -> self_collision_model.cpp:503-531), one thread per
// state, IEEE double, the reference's visit order (the result depends on it):
//   robotVoxelsCollisionDistance (:1386-1468): roots of the ROBOT's trees pushed in group order, popped from the back;
//     clearance = res * sqrt(d2(cell)) - (r + padding) (SphereCollisionDistance, collision_operations.h:81-89; 0 field
//     distance outside the grid); a sphere undercutting the bound d halves into it, d = max(0, 0.5 * clearance), and
//     splits, the child with the larger radius visited first; d == 0 ends the descent.
//   attached-body terms: "TODO: implement" in the reference, +infinity.
//   robotSpheresCollisionDistance (:1470-1487, 1512-1642): every checked robot pair contributes `return true` = 1.0.
__global__ void collision_distance_kernel(const DevModel* __restrict__ M, const uint16_t* __restrict__ df, GridParams G,
                                          double res, double padding, const double* __restrict__ q, int n,
                                          double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const double* qa = q + (size_t)i * M->dof;
    Xf T[MAX_LINKS];   // per-thread (local memory): every link transform, the descent visits trees in stack order
    for (int l = 0; l < M->n_links; ++l) {
        const int v = M->link_var[l];
        const double val = v >= 0 ? qa[v] : M->link_const[l];
        Xf J, P;
        joint_transform(M, l, val, J);
        const int p = M->link_parent[l];
        if (p < 0) {
#pragma unroll
            for (int k = 0; k < 12; ++k) P.m[k] = M->link_base[l][k];
        } else {
            P = T[p];
        }
        xf_mul(P, J, T[l]);
    }
    int stack[MAX_TREES + MAX_TREE_DEPTH];
    int sp = 0;
    for (int t = 0; t < M->n_robot_trees; ++t) {
        stack[sp++] = M->tree_root[t];
    }
    double d = __longlong_as_double(0x7FF0000000000000LL);   // +infinity
    while (sp > 0) {
        const int node = stack[--sp];
        double x, y, z;
        xf_point(T[M->node_link[node]], M->node_center[node][0], M->node_center[node][1], M->node_center[node][2], x, y, z);
        const int d2 = df_lookup(df, G, x, y, z);            // 0 outside the grid, like getMetricDistance
        const double dist = res * sqrt((double)d2);          // m_sqrt_table (distance_map.hpp:142-146)
        const double effective_radius = M->node_radius[node] + padding;
        const double obs_dist = dist - effective_radius;
        if (obs_dist >= d) {
            continue;
        }
        d = fmax(0.0, (1.0 - 0.5) * obs_dist);
        if (d == 0.0) {
            break;
        }
        const int left = M->node_left[node];
        if (left < 0) {
            continue;
        }
        const int right = M->node_right[node];
        if (M->node_radius[left] > M->node_radius[right]) {
            stack[sp++] = right;
            stack[sp++] = left;
        } else {
            stack[sp++] = left;
            stack[sp++] = right;
        }
    }
    if (M->n_robot_pairs > 0 && 1.0 < d) {
        d = 1.0;
    }
    out[i] = d;
}

// KDLRobotModel::checkJointLimits (kdl_robot_model.cpp:173-189, 210-235, 326-337) for one state
__device__ __forceinline__ bool joint_limits_ok(const DevModel* __restrict__ M, const double* q)
{
    const double PI = 3.14159265358979323846;
    const int dof = M->dof;
    bool good = true;
    for (int v = 0; v < dof; ++v) {
        if (M->var_min[v] > M->var_max[v]) {
            good = false;
        }
    }
    for (int v = 0; v < dof && good; ++v) {
        double a = q[v];
        const double a_min = M->var_min[v];
        const double a_max = M->var_min_norm[v]; // the reference passes normalize_angle(min) as a_max (:227-228)
        if (fabs(a) > 2.0 * PI) {
            a = fmod(a, 2.0 * PI);
        }
        while (a > a_max) {
            a -= 2.0 * PI;
        }
        while (a < a_min) {
            a += 2.0 * PI;
        }
        if (a < M->var_min[v] || a > M->var_max[v]) {
            good = false;
        }
    }
    return good;
}

__global__ void joint_limits_kernel(const DevModel* __restrict__ M, const double* __restrict__ q, int n,
                                    uint8_t* __restrict__ ok)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    ok[i] = joint_limits_ok(M, q + (size_t)i * M->dof) ? 1 : 0;
}

} // namespace smplgpu
