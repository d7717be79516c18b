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
// plus two per-row "level stamps": row_stamp (last level at which the row
// received frontier bits) and cand_stamp (last level at which the row or one of
// its eight (y,z) neighbours did), so a sweep touches only the 3x3 row
// neighbourhood of the wavefront.  Each level a warp scans 32 cand_stamps at a
// time (one coalesced 128-byte load), __ballot_sync/__ffs compacts the
// candidate rows, and the whole warp then expands one row: lanes = words, the
// nine neighbour rows are OR-ed, x-dilation is done with shifts and shuffles,
// and distances are written with one coalesced store per non-empty word.  One
// persistent cooperative kernel runs all levels with grid-wide barriers; frontier bitmaps stay L2 resident.
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
    uint32_t* front[2];
    // stamps are ping-ponged by level parity: level L writes [L&1] and reads [(L-1)&1],
    // so a row that is re-stamped during level L still reads as "active at L-1"
    uint32_t* row_stamp[2];   // [rows] last level (of that parity) at which the row received frontier bits
    uint32_t* cand_stamp[2];  // [rows] last level at which the row or one of its 8 neighbours did
    int* dist;             // [DZ*DY*DX]
    int* ctrl;             // [0] = levels run, [1..3] = rotating new-cell flags
};

constexpr uint32_t STAMP_NEVER = 0xFFFFFFFFu;

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
__global__ void bfs_walls_from_df_kernel(BfsGrid g, const uint16_t* __restrict__ df, int d2_wall_max,
                                         unsigned int* __restrict__ wall_count)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int cnt = 0;
    if (idx < g.rows * g.W) {
        const int row = idx / g.W, w = idx - row * g.W;
        const int z = row / g.DY, y = row - z * g.DY;
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            const int x = w * 32 + b;
            bool wall;
            if (x >= g.DX) {
                wall = true;
            } else if (x == 0 || x == g.DX - 1 || y == 0 || y == g.DY - 1 || z == 0 || z == g.DZ - 1) {
                wall = true;
            } else {
                // distance field is x-major / z-fastest
                const int d2 = df[((size_t)(x - 1) * g.ny + (y - 1)) * g.nz + (z - 1)];
                wall = d2 <= d2_wall_max;
                cnt += wall ? 1u : 0u;
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
        g.row_stamp[0][idx] = STAMP_NEVER;
        g.row_stamp[1][idx] = STAMP_NEVER;
        g.cand_stamp[0][idx] = STAMP_NEVER;
        g.cand_stamp[1][idx] = STAMP_NEVER;
    }
    if (idx < 4) {
        g.ctrl[idx] = 0;
    }
    if (idx >= g.rows * g.W) {
        return;
    }
    const uint32_t wbits = g.wall[idx];
    g.blocked[idx] = wbits;
    g.front[0][idx] = 0;
    g.front[1][idx] = 0;
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
    atomicOr(&g.front[0][word], bit);
    g.dist[(size_t)row * g.DX + px] = 0;
    g.row_stamp[0][row] = 0;
    for (int k = 0; k < 9; ++k) {
        g.cand_stamp[0][(pz + k / 3 - 1) * g.DY + (py + k % 3 - 1)] = 0;
    }
}

// OR of the frontier words `w` of the (up to nine) active neighbour rows of (y,z)
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

// Process one candidate row (whole warp).  Returns true when the row received new cells.
__device__ __forceinline__ bool bfs_expand_row(const BfsGrid& g, const uint32_t* __restrict__ fcur,
                                               uint32_t* __restrict__ fnext, int row, uint32_t level, int lane)
{
    const int z = row / g.DY, y = row - z * g.DY;
    const uint32_t prev = level - 1;
    // which of the nine neighbour rows carry frontier bits of the previous level
    uint32_t rs = STAMP_NEVER;
    if (lane < 9) {
        rs = __ldcg(&g.row_stamp[prev & 1][(z + lane / 3 - 1) * g.DY + (y + lane % 3 - 1)]);
    }
    const uint32_t active9 = __ballot_sync(0xffffffffu, rs == prev) & 0x1FFu;
    if (active9 == 0) {
        return false;
    }
    bool row_new = false;
    const int chunks = (g.W + 31) / 32;
    for (int c = 0; c < chunks; ++c) {
        const int w = c * 32 + lane;
        uint32_t m = 0;
        if (w < g.W) {
            m = gather9(fcur, g, y, z, w, active9);
        }
        uint32_t left = __shfl_up_sync(0xffffffffu, m, 1);
        uint32_t right = __shfl_down_sync(0xffffffffu, m, 1);
        if (lane == 0) {
            left = (w > 0) ? gather9(fcur, g, y, z, w - 1, active9) : 0;
        }
        if (lane == 31) {
            right = (w + 1 < g.W) ? gather9(fcur, g, y, z, w + 1, active9) : 0;
        }
        uint32_t fresh = 0;
        if (w < g.W) {
            const uint32_t dil = m | (m << 1) | (m >> 1) | (left >> 31) | (right << 31);
            const size_t idx = (size_t)row * g.W + w;
            const uint32_t blk = __ldcg(&g.blocked[idx]);
            fresh = dil & ~blk;
            if (fresh) {
                g.blocked[idx] = blk | fresh;
            }
            // a stamped row must have every word current; unstamped rows are never read
            fnext[idx] = fresh;
        }
        // distances: one coalesced 128-byte store per non-empty word
        uint32_t nz = __ballot_sync(0xffffffffu, fresh != 0);
        row_new |= nz != 0;
        while (nz) {
            const int j = __ffs(nz) - 1;
            nz &= nz - 1;
            const uint32_t wj = __shfl_sync(0xffffffffu, fresh, j);
            if ((wj >> lane) & 1u) {
                g.dist[(size_t)row * g.DX + (size_t)(c * 32 + j) * 32 + lane] = (int)level;
            }
        }
    }
    if (row_new && lane < 9) {
        // mark this row and its eight neighbours as candidates for the next level
        g.cand_stamp[level & 1][(z + lane / 3 - 1) * g.DY + (y + lane % 3 - 1)] = level;
        if (lane == 0) {
            g.row_stamp[level & 1][row] = level;
        }
    }
    return row_new;
}

// All levels in one cooperative launch.  ctrl[1..3] are rotating "new cells at
// level L" flags: flag[L%3] is set during level L, read after the barrier, and
// flag[(L+1)%3] is cleared during level L (its last readers ran before the
// previous barrier), so one grid barrier per level suffices.
__global__ void __launch_bounds__(1024)
bfs_levels_kernel(BfsGrid g, int max_levels)
{
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int gwarp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;

    uint32_t level = 1;
    for (; level <= (uint32_t)max_levels; ++level) {
        const uint32_t* __restrict__ fcur = g.front[(level - 1) & 1];
        uint32_t* __restrict__ fnext = g.front[level & 1];
        const uint32_t prev = level - 1;
        int* newflag = &g.ctrl[1 + level % 3];
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            g.ctrl[1 + (level + 1) % 3] = 0;
        }
        bool any_new = false;
        // scan: lane <-> row, 32 rows per warp step; ballot-compact the candidates
        for (int base = gwarp * 32; base < g.rows; base += nwarps * 32) {
            const int r = base + lane;
            uint32_t st = STAMP_NEVER;
            if (r < g.rows) {
                st = __ldcg(&g.cand_stamp[prev & 1][r]);
            }
            uint32_t cand = __ballot_sync(0xffffffffu, st == prev);
            while (cand) {
                const int j = __ffs(cand) - 1;
                cand &= cand - 1;
                const int row = base + j;
                const int z = row / g.DY, y = row - z * g.DY;
                if (z == 0 || z == g.DZ - 1 || y == 0 || y == g.DY - 1) {
                    continue; // border shell rows are all wall
                }
                any_new |= bfs_expand_row(g, fcur, fnext, row, level, lane);
            }
        }
        if (any_new && lane == 0) {
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
