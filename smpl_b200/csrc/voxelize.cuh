// Scene ingest on the device: surface voxelisation of triangle meshes, the step between
// a collision object and the distance field in the reference:
//   geometry::VoxelizeMesh / VoxelizeBox       smpl/src/geometry/voxelize.cpp:673-736, 962-1054
//   geometry::VoxelizeTriangle                 smpl/include/smpl/geometry/detail/voxelize.hpp:45-181
//   Distance (point / capsule)                 voxelize.cpp:626-649
//   PivotDiscretizer / HalfResDiscretizer      smpl/include/smpl/geometry/discretize.h:41-112
//   ExtractVoxels                              voxelize.cpp:206-222
//   OccupancyGrid::addPointsToField -> DistanceMap::addPointsToMap   occupancy_grid.cpp:357-382, distance_map.hpp:305-328
//
// The reference walks the cells of every triangle's bounding box one after the other; a cell's verdict depends only
// on (triangle, cell), and the grid is the OR over triangles, so the work is flattened into (triangle, cell) items:
// one thread per item, consecutive lanes on consecutive z cells of one triangle.  All arithmetic is IEEE double in
// the reference's operation order (the translation unit is compiled with --fmad=false; sqrt and division are
// correctly rounded), so the voxel set is the reference's bit for bit.
#pragma once

#include <stdint.h>

#include "model.cuh"

namespace smplgpu {

struct VoxDisc
{
    int half_res;        // 1: HalfResVoxelGrid (cells centred on (i + 1/2) res), 0: PivotVoxelGrid (on pivot + i res)
    double res;
    double pivot[3];
};

// per-triangle constants of VoxelizeTriangle (voxelize.hpp:52-123)
struct TriSetup
{
    double p1[3], p2[3], p3[3];
    double n[3], e1[3], e2[3], e3[3];
    double d, t, d1, d2, d3, rc2;
    int mn[3];           // first cell of the bounding box
    int ext[3];          // cells per axis (0 for a degenerate triangle)
};

__device__ __forceinline__ int vox_discretize(const VoxDisc& D, int axis, double d)
{
    if (D.half_res) {
        return (d >= 0) ? __double2int_rz(d / D.res) : (__double2int_rz(d / D.res) - 1);
    }
    return (int)floor((d - D.pivot[axis]) / D.res + 0.5);
}

__device__ __forceinline__ double vox_continuize(const VoxDisc& D, int axis, int i)
{
    if (D.half_res) {
        return (double)i * D.res + 0.5 * D.res;
    }
    return D.pivot[axis] + i * D.res;
}

__device__ __forceinline__ double v3_dot(const double* a, const double* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

__device__ __forceinline__ void v3_cross(const double* a, const double* b, double* r)
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

// Eigen 3.3 normalize(): divide by the norm when the squared norm is positive
__device__ __forceinline__ void v3_normalize(double* a)
{
    const double z = v3_dot(a, a);
    if (z > 0.0) {
        const double s = sqrt(z);
        a[0] /= s; a[1] /= s; a[2] /= s;
    }
}

__global__ void vox_setup_kernel(const double* __restrict__ vertices, const int* __restrict__ tris, int nt,
                                 VoxDisc D, TriSetup* __restrict__ out, unsigned long long* __restrict__ counts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) {
        return;
    }
    TriSetup S;
    const double* a = vertices + 3 * (size_t)tris[3 * i];
    const double* b = vertices + 3 * (size_t)tris[3 * i + 1];
    const double* c = vertices + 3 * (size_t)tris[3 * i + 2];
    double u[3], v[3], w[3], ac[3], tmp[3];
    for (int k = 0; k < 3; ++k) {
        S.p1[k] = a[k]; S.p2[k] = b[k]; S.p3[k] = c[k];
        u[k] = b[k] - a[k];          // p2 - p1
        v[k] = c[k] - b[k];          // p3 - p2
        w[k] = a[k] - c[k];          // p1 - p3
        ac[k] = c[k] - a[k];         // p3 - p1
    }
    // colinearity: ((p2 - p1) x (p3 - p1)).norm() == 0  (a norm is never negative, so the ccw swap never happens)
    v3_cross(u, ac, tmp);
    const double det = sqrt(v3_dot(tmp, tmp));
    S.ext[0] = S.ext[1] = S.ext[2] = 0;
    counts[i] = 0;
    if (det == 0) {
        out[i] = S;
        return;
    }
    const double rc = sqrt(3.0) * 0.5 * D.res;
    S.rc2 = rc * rc;
    v3_cross(u, v, S.n);
    v3_normalize(S.n);
    const double k = 0.5774;
    double ca = 0.0;
    for (int m = 0; m < 8; ++m) {
        const double corner[3] = { (m & 4) ? k : -k, (m & 2) ? k : -k, (m & 1) ? k : -k };
        const double cv = v3_dot(corner, S.n);
        ca = (m == 0) ? cv : fmax(ca, cv);
    }
    S.t = rc * ca;
    S.d = -v3_dot(S.n, S.p1);
    v3_cross(u, S.n, S.e1);
    v3_cross(v, S.n, S.e2);
    v3_cross(w, S.n, S.e3);
    for (int q = 0; q < 3; ++q) {
        S.e1[q] = -S.e1[q]; S.e2[q] = -S.e2[q]; S.e3[q] = -S.e3[q];
    }
    v3_normalize(S.e1);
    v3_normalize(S.e2);
    v3_normalize(S.e3);
    S.d1 = -v3_dot(S.e1, S.p1);
    S.d2 = -v3_dot(S.e2, S.p2);
    S.d3 = -v3_dot(S.e3, S.p3);
    unsigned long long cells = 1;
    for (int q = 0; q < 3; ++q) {
        const double lo = fmin(a[q], fmin(b[q], c[q]));
        const double hi = fmax(a[q], fmax(b[q], c[q]));
        S.mn[q] = vox_discretize(D, q, lo);
        S.ext[q] = vox_discretize(D, q, hi) - S.mn[q] + 1;
        cells *= (unsigned long long)max(S.ext[q], 0);
    }
    counts[i] = cells;
    out[i] = S;
}

// voxelize.cpp:626-649
__device__ __forceinline__ bool vox_on_capsule(const double* p, const double* q, double radius_sqrd, const double* x)
{
    double pq[3], px[3];
    for (int k = 0; k < 3; ++k) {
        pq[k] = q[k] - p[k];
        px[k] = x[k] - p[k];
    }
    const double d = v3_dot(px, pq);
    const double l2 = v3_dot(pq, pq);
    if (d < 0.0 || d > l2) {
        return false;
    }
    const double dsq = v3_dot(px, px) - (d * d) / l2;
    return !(dsq > radius_sqrd);   // the reference returns dsq (never -1.0 here) and compares the result with -1.0
}

__device__ __forceinline__ double vox_sign(double v) { return (v == 0) ? 0.0 : ((v > 0) ? 1.0 : -1.0); }

// does triangle S fill the voxel centred on x? (voxelize.hpp:125-176)
__device__ __forceinline__ bool vox_cell_filled(const TriSetup& S, const double* x)
{
    double dx[3];
    for (int k = 0; k < 3; ++k) dx[k] = x[k] - S.p1[k];
    if (v3_dot(dx, dx) <= S.rc2) return true;
    for (int k = 0; k < 3; ++k) dx[k] = x[k] - S.p2[k];
    if (v3_dot(dx, dx) <= S.rc2) return true;
    for (int k = 0; k < 3; ++k) dx[k] = x[k] - S.p3[k];
    if (v3_dot(dx, dx) <= S.rc2) return true;
    // the reference asks for the edges (p1, p3), (p2, p3), (p3, p1): p1-p2 is never tested
    if (vox_on_capsule(S.p1, S.p3, S.rc2, x) || vox_on_capsule(S.p2, S.p3, S.rc2, x) || vox_on_capsule(S.p3, S.p1, S.rc2, x)) {
        return true;
    }
    const double nx = v3_dot(S.n, x);
    if (vox_sign(nx + (S.d + S.t)) == vox_sign(nx + (S.d - S.t))) {
        return false;
    }
    return v3_dot(S.e1, x) + S.d1 > 0.0 && v3_dot(S.e2, x) + S.d2 > 0.0 && v3_dot(S.e3, x) + S.d3 > 0.0;
}

// One thread per (triangle, cell) item.  offsets[nt + 1] = exclusive prefix of the per-triangle cell counts.
// mode 0: set the cell's bit in a bitmap over the mesh's own voxel grid (gmin / gext, x-major, z fastest);
// mode 1: addPointsToField -- the voxel centre goes through DistanceMap::worldToGrid and marks occ[] when inside.
__global__ void vox_cells_kernel(const TriSetup* __restrict__ setup, const unsigned long long* __restrict__ offsets, int nt,
                                 unsigned long long total, VoxDisc D, int mode, int3 gmin, int3 gext,
                                 unsigned int* __restrict__ bits, GridParams G, uint8_t* __restrict__ occ)
{
    for (unsigned long long item = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; item < total;
         item += (unsigned long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = nt;   // last triangle whose offset is <= item
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (offsets[mid] <= item) lo = mid; else hi = mid;
        }
        const TriSetup& S = setup[lo];
        unsigned long long r = item - offsets[lo];
        const int cz = (int)(r % (unsigned)S.ext[2]);
        r /= (unsigned)S.ext[2];
        const int cy = (int)(r % (unsigned)S.ext[1]);
        const int cx = (int)(r / (unsigned)S.ext[1]);
        const int gx = S.mn[0] + cx, gy = S.mn[1] + cy, gz = S.mn[2] + cz;
        const double x[3] = { vox_continuize(D, 0, gx), vox_continuize(D, 1, gy), vox_continuize(D, 2, gz) };
        if (!vox_cell_filled(S, x)) {
            continue;
        }
        if (mode == 0) {
            if ((unsigned)(gx - gmin.x) >= (unsigned)gext.x || (unsigned)(gy - gmin.y) >= (unsigned)gext.y ||
                (unsigned)(gz - gmin.z) >= (unsigned)gext.z) {
                continue;   // min + (max - min) rounded below max: the reference would index outside its grid
            }
            const unsigned long long idx = ((unsigned long long)(gx - gmin.x) * gext.y + (unsigned)(gy - gmin.y)) * gext.z + (unsigned)(gz - gmin.z);
            atomicOr(&bits[idx >> 5], 1u << (idx & 31));
        } else {
            const int fx = __double2int_rz(G.inv_res * (x[0] - G.ox) + 0.5) - 1;
            const int fy = __double2int_rz(G.inv_res * (x[1] - G.oy) + 0.5) - 1;
            const int fz = __double2int_rz(G.inv_res * (x[2] - G.oz) + 0.5) - 1;
            if ((unsigned)fx < (unsigned)G.nx && (unsigned)fy < (unsigned)G.ny && (unsigned)fz < (unsigned)G.nz) {
                occ[((size_t)fx * G.ny + fy) * G.nz + fz] = 1;
            }
        }
    }
}

// ---- ExtractVoxels: ordered compaction of the bitmap (memory order = x, then y, then z) ----
constexpr int VOX_SCAN_THREADS = 256;

// per-word popcounts -> exclusive prefix inside each block of 256 words + the block's total
__global__ void vox_count_kernel(const unsigned int* __restrict__ bits, size_t n_words, unsigned int* __restrict__ prefix,
                                 unsigned int* __restrict__ block_total)
{
    __shared__ unsigned int sh[VOX_SCAN_THREADS];
    const size_t w = (size_t)blockIdx.x * VOX_SCAN_THREADS + threadIdx.x;
    const unsigned int c = w < n_words ? __popc(bits[w]) : 0u;
    sh[threadIdx.x] = c;
    __syncthreads();
    for (int off = 1; off < VOX_SCAN_THREADS; off <<= 1) {
        const unsigned int add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
        __syncthreads();
        sh[threadIdx.x] += add;
        __syncthreads();
    }
    if (w < n_words) {
        prefix[w] = sh[threadIdx.x] - c;
    }
    if (threadIdx.x == VOX_SCAN_THREADS - 1) {
        block_total[blockIdx.x] = sh[threadIdx.x];
    }
}

// one block: exclusive prefix over the block totals (in place), grand total in *total
__global__ void vox_scan_blocks_kernel(unsigned int* __restrict__ block_total, int n_blocks, unsigned int* __restrict__ total)
{
    __shared__ unsigned int sh[VOX_SCAN_THREADS];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += VOX_SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const unsigned int c = i < n_blocks ? block_total[i] : 0u;
        sh[threadIdx.x] = c;
        __syncthreads();
        for (int off = 1; off < VOX_SCAN_THREADS; off <<= 1) {
            const unsigned int add = threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
            __syncthreads();
            sh[threadIdx.x] += add;
            __syncthreads();
        }
        if (i < n_blocks) {
            block_total[i] = carry + sh[threadIdx.x] - c;
        }
        __syncthreads();
        if (threadIdx.x == VOX_SCAN_THREADS - 1) carry += sh[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void vox_extract_kernel(const unsigned int* __restrict__ bits, size_t n_words, const unsigned int* __restrict__ prefix,
                                   const unsigned int* __restrict__ block_base, VoxDisc D, int3 gmin, int3 gext,
                                   double* __restrict__ out, unsigned int max_voxels)
{
    const size_t w = (size_t)blockIdx.x * VOX_SCAN_THREADS + threadIdx.x;
    if (w >= n_words) {
        return;
    }
    unsigned int word = bits[w];
    unsigned int k = block_base[blockIdx.x] + prefix[w];
    while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        if (k < max_voxels) {
            unsigned long long idx = (unsigned long long)w * 32 + b;
            const int cz = (int)(idx % (unsigned)gext.z);
            idx /= (unsigned)gext.z;
            const int cy = (int)(idx % (unsigned)gext.y);
            const int cx = (int)(idx / (unsigned)gext.y);
            out[3 * (size_t)k] = vox_continuize(D, 0, gmin.x + cx);
            out[3 * (size_t)k + 1] = vox_continuize(D, 1, gmin.y + cy);
            out[3 * (size_t)k + 2] = vox_continuize(D, 2, gmin.z + cz);
        }
        ++k;
    }
}

} // namespace smplgpu
