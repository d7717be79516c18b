// Kernel (4): BFS_3D as a level-synchronous wavefront over bit-packed grids.
//
// Reference (file:line under dyouakim/smpl): smpl/src/bfs3d.cpp:40-111 (grid
// layout, WALL border), :156-201 (run: reset non-walls, seed), :501-547
// (search: 26-connected unit-cost FIFO expansion); smpl/include/smpl/bfs3d/bfs3d.h.
// With unit edge costs every level-synchronous order yields the FIFO search's
// distances, so the int32 result is bit-identical.
//
// Layout.  Padded dims DX=nx+2, DY=ny+2, DZ=nz+2, node = (z*DY + y)*DX + x as in
// the reference.  One bit per cell in rows of W 32-bit words (W*32 >= DX, W a
// multiple of 4 so a row is a whole number of 128-bit words):
//   wall     walls incl. the border shell (persistent across runs)
//   blocked  wall | discovered
//   front[2] cells discovered at the previous / current level (ping-pong)
// plus a per-row candidate word (level, mask of the nine (y,z) neighbour rows
// that received cells at that level), so a sweep touches only the 3x3 row
// neighbourhood of the wavefront.  Each level a warp scans 32 candidate words,
// __ballot_sync/__ffs compacts the candidate rows, and each half-warp expands
// one row: lanes = words, the active neighbour rows are OR-ed, x-dilation is
// done with shifts and shuffles, and distances are written with coalesced
// stores per non-empty word.  One persistent cooperative kernel runs all levels with grid-wide barriers; frontier bitmaps stay L2 resident.
#pragma once

#include <cooperative_groups.h>
#include <stdint.h>

namespace smplgpu {

namespace cg = cooperative_groups;

struct BfsGrid
{
    int nx, ny, nz;
    int DX, DY, DZ;
    int W;                 // words per row
    int rows;              // DY * DZ
    uint32_t* wall;
    uint32_t* blocked;
    uint32_t* front0;
    uint32_t* front1;
    // candidate words, ping-ponged by level parity: level L writes [L&1] and reads [(L-1)&1]
    uint32_t* cand0;          // [rows] (level << 9) | mask of neighbour rows that got cells at `level`
    uint32_t* cand1;
    int* dist;             // [DZ*DY*DX]
    int* ctrl;             // [0] = levels run, [1..3] = rotating new-cell flags
};


// wall bitmap from one byte per cell (x fastest, unpadded); border shell = wall
__global__ void bfs_walls_from_bytes_kernel(BfsGrid g, const uint8_t* __restrict__ walls)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.rows * g.W) {
        return;
    }
    const int row = idx / g.W, w = idx - row * g.W;
    const int z = row / g.DY, y = row - z * g.DY;
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
        const int x = w * 32 + b;
        bool wall;
        if (x >= g.DX) {
            wall = true; // padding beyond the row end never opens
        } else if (x == 0 || x == g.DX - 1 || y == 0 || y == g.DY - 1 || z == 0 || z == g.DZ - 1) {
            wall = true;
        } else {
            wall = walls[((size_t)(z - 1) * g.ny + (y - 1)) * g.nx + (x - 1)] != 0;
        }
        bits |= (wall ? 1u : 0u) << b;
    }
    g.wall[idx] = bits;
}

// BfsHeuristic::syncGridAndBfs: wall iff d2(cell) <= d2_wall_max  (<=> res*sqrt(d2) <= radius)
// slot_dz: padded z extent of one slot when several BFS_3D instances are stacked along z
// (== g.DZ for a single grid); every slot gets its own border shell.
__global__ void bfs_walls_from_df_kernel(BfsGrid g, const uint16_t* __restrict__ df, int d2_wall_max,
                                         int slot_dz, unsigned int* __restrict__ wall_count)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int cnt = 0;
    if (idx < g.rows * g.W) {
        const int row = idx / g.W, w = idx - row * g.W;
        const int zt = row / g.DY, y = row - zt * g.DY;
        const int z = zt % slot_dz;           // z within the slot
        const int nz = slot_dz - 2;
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int x = w * 32 + b;
            bool wall;
            if (x >= g.DX) {
                wall = true;
            } else if (x == 0 || x == g.DX - 1 || y == 0 || y == g.DY - 1 || z == 0 || z == slot_dz - 1) {
                wall = true;
            } else {
                // distance field is x-major / z-fastest
                const int d2 = df[((size_t)(x - 1) * g.ny + (y - 1)) * nz + (z - 1)];
                wall = d2 <= d2_wall_max;
                cnt += (wall && zt < slot_dz) ? 1u : 0u;   // count the first slot only
            }
            bits |= (wall ? 1u : 0u) << b;
        }
        g.wall[idx] = bits;
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(wall_count, cnt);
    }
}

// BFS_3D::run reset: non-walls -> UNDISCOVERED, blocked = wall, stamps cleared
__global__ void bfs_reset_kernel(BfsGrid g)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < g.rows) {
        g.cand0[idx] = 0;
        g.cand1[idx] = 0;
    }
    if (idx < 4) {
        g.ctrl[idx] = 0;
    }
    if (idx >= g.rows * g.W) {
        return;
    }
    const uint32_t wbits = g.wall[idx];
    g.blocked[idx] = wbits;
    g.front0[idx] = 0;
    g.front1[idx] = 0;
    const int row = idx / g.W, w = idx - row * g.W;
    int* d = g.dist + (size_t)row * g.DX + w * 32;
    const int xmax = min(32, g.DX - w * 32);
    for (int b = 0; b < xmax; ++b) {
        d[b] = ((wbits >> b) & 1u) ? 0x7FFFFFFF : -1;
    }
}

// seeds: grid[origin] = 0 even when origin is a wall (bfs3d.cpp:181-187), which
// permanently turns that wall cell into a free cell for later runs too.
__global__ void bfs_seed_kernel(BfsGrid g, const int* __restrict__ seeds, int n_seeds, int* __restrict__ n_in_bounds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seeds) {
        return;
    }
    const int x = seeds[3 * i], y = seeds[3 * i + 1], z = seeds[3 * i + 2];
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) {
        return;
    }
    atomicAdd(n_in_bounds, 1);
    const int px = x + 1, py = y + 1, pz = z + 1;
    const int row = pz * g.DY + py;
    const size_t word = (size_t)row * g.W + (px >> 5);
    const uint32_t bit = 1u << (px & 31);
    atomicAnd(&g.wall[word], ~bit);
    atomicOr(&g.blocked[word], bit);
    atomicOr(&g.front0[word], bit);
    g.dist[(size_t)row * g.DX + px] = 0;
    // level 0 published to the nine neighbour rows: the seed row is neighbour k of (pz - dz, py - dy)
    for (int k = 0; k < 9; ++k) {
        atomicOr(&g.cand0[(pz - (k / 3 - 1)) * g.DY + (py - (k % 3 - 1))], 1u << k);
    }
}

// OR of the frontier words `w` of the active neighbour rows of (y,z); bit k of
// `active9` <=> neighbour row (z + k/3 - 1, y + k%3 - 1) has frontier bits
__device__ __forceinline__ uint32_t gather9(const uint32_t* __restrict__ fr, const BfsGrid& g,
                                            int y, int z, int w, uint32_t active9)
{
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (active9 & (1u << k)) {
            const int nz_ = z + k / 3 - 1, ny_ = y + k % 3 - 1;
            m |= __ldcg(&fr[(size_t)(nz_ * g.DY + ny_) * g.W + w]);
        }
    }
    return m;
}

constexpr int BFS_THREADS = 512;

// All levels in one cooperative launch.
//
// cand[p][row] = (level << 9) | mask9: bit k of mask9 says that neighbour row k of
// `row` received frontier cells at `level` (parity p = level & 1).  A row that gets
// new cells at level L publishes itself to its nine (y,z) neighbours with
// atomicMax (moves the word to level L, clearing older mask bits) + atomicOr.
// Level L+1 scans cand[L & 1]: rows are interleaved over warps (row = warp + j *
// nwarps) so the wavefront's rows spread evenly, each lane loads one word,
// __ballot_sync/__ffs compacts the candidates, and the warp then expands two
// candidate rows at a time, one per half-warp (lanes = 32-bit words of the row).
//
// ctrl[1..3] are rotating "new cells at level L" flags: flag[L%3] is set during
// level L and read after the barrier; flag[(L+1)%3] is cleared during level L (its
// last readers ran before the previous barrier): one grid barrier per level.
__global__ void __launch_bounds__(BFS_THREADS, 3)
bfs_levels_kernel(const __grid_constant__ BfsGrid g, int max_levels)
{
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int half = lane >> 4, sl = lane & 15;
    const int warps_per_block = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    const int chunks = (g.W + 15) / 16;

    uint32_t level = 1;
    for (; level <= (uint32_t)max_levels; ++level) {
        const bool odd = level & 1;
        const uint32_t* __restrict__ fcur = odd ? g.front0 : g.front1;
        uint32_t* __restrict__ fnext = odd ? g.front1 : g.front0;
        const uint32_t prev = level - 1;
        const uint32_t* __restrict__ cand_in = odd ? g.cand0 : g.cand1;
        uint32_t* __restrict__ cand_out = odd ? g.cand1 : g.cand0;
        int* newflag = &g.ctrl[1 + level % 3];
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            g.ctrl[1 + (level + 1) % 3] = 0;
        }
        bool any_new = false;
        for (int base = gwarp; base < g.rows; base += nwarps * 32) {
            const int myrow = base + lane * nwarps;
            uint32_t word = 0;
            if (myrow < g.rows) {
                word = __ldcg(&cand_in[myrow]);
            }
            uint32_t cmask = __ballot_sync(0xffffffffu, (word >> 9) == prev && (word & 0x1FFu) != 0);
            while (cmask) {
                // two candidates per pass, one per half-warp
                const int j0 = __ffs(cmask) - 1;
                cmask &= cmask - 1;
                int j1 = -1;
                if (cmask) {
                    j1 = __ffs(cmask) - 1;
                    cmask &= cmask - 1;
                }
                const int j = half ? j1 : j0;
                const uint32_t w9 = __shfl_sync(0xffffffffu, word, j < 0 ? 0 : j);
                int row = j < 0 ? -1 : base + j * nwarps;
                int z = 0, y = 0;
                if (row >= 0) {
                    z = row / g.DY;
                    y = row - z * g.DY;
                    if (z == 0 || z == g.DZ - 1 || y == 0 || y == g.DY - 1) {
                        row = -1; // border shell rows are all wall
                    }
                }
                const uint32_t active9 = row >= 0 ? (w9 & 0x1FFu) : 0u;
                bool row_new = false;
                for (int c = 0; c < chunks; ++c) {
                    const int w = c * 16 + sl;
                    const bool on = row >= 0 && w < g.W;
                    uint32_t m = 0, blk = 0xFFFFFFFFu;
                    size_t idx = 0;
                    if (on) {
                        idx = (size_t)row * g.W + w;
                        blk = __ldcg(&g.blocked[idx]);
                        m = gather9(fcur, g, y, z, w, active9);
                    }
                    uint32_t left = __shfl_up_sync(0xffffffffu, m, 1, 16);
                    uint32_t right = __shfl_down_sync(0xffffffffu, m, 1, 16);
                    if (sl == 0) {
                        left = (on && w > 0) ? gather9(fcur, g, y, z, w - 1, active9) : 0;
                    }
                    if (sl == 15) {
                        right = (on && w + 1 < g.W) ? gather9(fcur, g, y, z, w + 1, active9) : 0;
                    }
                    uint32_t fresh = 0;
                    if (on) {
                        const uint32_t dil = m | (m << 1) | (m >> 1) | (left >> 31) | (right << 31);
                        fresh = dil & ~blk;
                        if (fresh) {
                            g.blocked[idx] = blk | fresh;
                        }
                        fnext[idx] = fresh; // a published row has every word current
                    }
                    // distances: per non-empty word, two coalesced 64-byte stores per half-warp
                    uint32_t nz = (__ballot_sync(0xffffffffu, fresh != 0) >> (half * 16)) & 0xFFFFu;
                    row_new |= nz != 0;
                    const uint32_t hmask = half ? 0xFFFF0000u : 0x0000FFFFu; // the halves loop independently
                    while (nz) {
                        const int k = __ffs(nz) - 1;
                        nz &= nz - 1;
                        const uint32_t wk = __shfl_sync(hmask, fresh, half * 16 + k);
                        int* d = g.dist + (size_t)row * g.DX + (size_t)(c * 16 + k) * 32;
                        if ((wk >> sl) & 1u) {
                            d[sl] = (int)level;
                        }
                        if ((wk >> (sl + 16)) & 1u) {
                            d[sl + 16] = (int)level;
                        }
                    }
                }
                if (row_new) {
                    if (sl < 9) {
                        // this row is neighbour k = sl of row (z - dz, y - dy)
                        const int dz = sl / 3 - 1, dy = sl % 3 - 1;
                        uint32_t* cw = &cand_out[(z - dz) * g.DY + (y - dy)];
                        atomicMax(cw, level << 9);
                        atomicOr(cw, 1u << sl);
                    }
                    any_new = true;
                }
            }
        }
        if (__any_sync(0xffffffffu, any_new) && lane == 0) {
            *newflag = 1;
        }
        grid.sync();
        if (!__ldcg(newflag)) {
            break;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        g.ctrl[0] = (int)level;
    }
}

// BFS_3D::getDistance for a list of cells
__global__ void bfs_gather_kernel(BfsGrid g, const int* __restrict__ cells, int n, int* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const int x = cells[3 * i], y = cells[3 * i + 1], z = cells[3 * i + 2];
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) {
        out[i] = -2;
        return;
    }
    out[i] = g.dist[((size_t)(z + 1) * g.DY + (y + 1)) * g.DX + (x + 1)];
}

} // namespace smplgpu
