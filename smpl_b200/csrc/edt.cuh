// Distance-field construction on the device ("next" row f1): exact squared
// Euclidean distance, in cells, to the nearest occupied cell OR border cell,
// capped at dmax^2 -- the quantity DistanceMap<EuclidDistanceMap> maintains
// incrementally on the host (smpl/include/smpl/distance_map/detail/
// distance_map.hpp:111-180 border init, :305-328 addPointsToMap, :728-762
// propagate; smpl/src/distance_map/euclid_distance_map.cpp:49-56).
// Separable three-pass transform (z, then y, then x), each pass bounded to a
// +-dmax window.  Layout: x-major / z-fastest, unpadded, like Grid3.
#pragma once

#include <stdint.h>

namespace smplgpu {

__global__ void edt_scatter_kernel(const int* __restrict__ cells, int n, int nx, int ny, int nz, uint8_t* __restrict__ occ)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const int x = cells[3 * i], y = cells[3 * i + 1], z = cells[3 * i + 2];
    if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) {
        return; // DistanceMap::addPointsToMap ignores points outside the map
    }
    occ[((size_t)x * ny + y) * nz + z] = 1;
}

// value = 1: cells enter the obstacle set, value = 0: they leave it
__global__ void edt_scatter_value_kernel(const int* __restrict__ cells, int n, int nx, int ny, int nz, uint8_t value,
                                         uint8_t* __restrict__ occ)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const int x = cells[3 * i], y = cells[3 * i + 1], z = cells[3 * i + 2];
    if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) {
        return;
    }
    occ[((size_t)x * ny + y) * nz + z] = value;
}

// the obstacle set of a field: the cells at distance 0 (the border the reference treats as obstacles lies outside)
__global__ void edt_occ_from_field_kernel(const uint16_t* __restrict__ d2, size_t cells, uint8_t* __restrict__ occ)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cells) {
        occ[i] = d2[i] == 0 ? 1 : 0;
    }
}

// pass 1: 1-D distance along z to the nearest occupied cell or border (z = -1, z = nz)
__global__ void edt_pass_z_kernel(const uint8_t* __restrict__ occ, int nx, int ny, int nz, int dmax,
                                  uint16_t* __restrict__ g)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= nx * ny) {
        return;
    }
    const uint8_t* o = occ + (size_t)col * nz;
    uint16_t* out = g + (size_t)col * nz;
    int d = 0; // border cell at z = -1
    for (int z = 0; z < nz; ++z) {
        d = o[z] ? 0 : min(d + 1, dmax + 1);
        out[z] = (uint16_t)d;
    }
    d = 0; // border cell at z = nz
    for (int z = nz - 1; z >= 0; --z) {
        d = o[z] ? 0 : min(d + 1, dmax + 1);
        out[z] = (uint16_t)min((int)out[z], d);
    }
}

// pass 2/3: out(p) = min over q along the axis (border cells at -1 and n count
// with value 0) of in(q) + (p-q)^2, capped.  `first` selects squaring of the
// pass-1 1-D distances.
__global__ void edt_pass_axis_kernel(const uint16_t* __restrict__ in, int nx, int ny, int nz, int axis, int dmax,
                                     int cap, bool first, uint16_t* __restrict__ out)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)nx * ny * nz;
    if (idx >= total) {
        return;
    }
    const int y = (int)((idx / nz) % ny);
    const int x = (int)(idx / ((size_t)nz * ny));
    const int p = axis == 1 ? y : x;
    const int n = axis == 1 ? ny : nx;
    const size_t stride = axis == 1 ? (size_t)nz : (size_t)nz * ny;
    const size_t base = idx - (size_t)p * stride;
    int best = cap;
    // border cells
    best = min(best, (p + 1) * (p + 1));
    best = min(best, (n - p) * (n - p));
    const int lo = max(0, p - dmax), hi = min(n - 1, p + dmax);
    for (int q = lo; q <= hi; ++q) {
        int v = in[base + (size_t)q * stride];
        if (first) {
            v = v * v;
        }
        const int dq = p - q;
        best = min(best, v + dq * dq);
    }
    out[idx] = (uint16_t)min(best, cap);
}

} // namespace smplgpu
