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
    uint32_t* cand0;          // [rows] neighbour-row mask (bits 0-8) | published words mod 16 (bits 9-24)
    uint32_t* cand1;
    int* dist;             // [DZ*DY*DX]
    int* ctrl;             // [16]: [0] = levels run, [4..5] = 64-bit grid barrier word (arrivals), [8..11] = per-level
                           // "a block found new cells" flags of bfs_levels_kernel (index level & 3)
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
// slot_mask (optional): only slots with a non-zero entry are rewritten.
__global__ void bfs_walls_from_df_kernel(BfsGrid g, const uint16_t* __restrict__ df, int d2_wall_max,
                                         int slot_dz, unsigned int* __restrict__ wall_count,
                                         const uint8_t* __restrict__ slot_mask)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int cnt = 0;
    bool on = idx < g.rows * g.W;
    if (on && slot_mask != nullptr) {
        on = slot_mask[(idx / g.W / g.DY) / slot_dz] != 0;
    }
    if (on) {
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

// BFS_3D::run reset: non-walls -> UNDISCOVERED, blocked = wall, stamps cleared.
// slot_mask (optional): only the masked slots of a stacked bank are reset (candidate stamps always are).
// One warp per bitmap word: lane b writes the distance of bit b, so a word's 32 distances go out as one
// coalesced 128-byte store.
__global__ void bfs_reset_kernel(BfsGrid g, const uint8_t* __restrict__ slot_mask, int slot_dz)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < g.rows) {
        g.cand0[tid] = 0;
        g.cand1[tid] = 0;
    }
    if (tid < 16) {
        g.ctrl[tid] = 0;
    }
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = g.rows * g.W;
    for (int base = (tid >> 5) * 32; base < total; base += nwarps * 32) {
        // lane l owns word base + l for the bitmap stores
        const int idx = base + lane;
        uint32_t wbits = 0;
        bool on = idx < total;
        if (on && slot_mask != nullptr) {
            on = slot_mask[(idx / g.W / g.DY) / slot_dz] != 0;
        }
        if (on) {
            wbits = g.wall[idx];
            g.blocked[idx] = wbits;
            g.front0[idx] = 0;
            g.front1[idx] = 0;
        }
        // distances: the warp walks its 32 words together
        const uint32_t live = __ballot_sync(0xffffffffu, on);
        for (int k = 0; k < 32; ++k) {
            if (!((live >> k) & 1u)) {
                continue;
            }
            const uint32_t wk = __shfl_sync(0xffffffffu, wbits, k);
            const int id = base + k;
            const int row = id / g.W, w = id - row * g.W;
            const int x = w * 32 + lane;
            if (x < g.DX) {
                g.dist[(size_t)row * g.DX + x] = ((wk >> lane) & 1u) ? 0x7FFFFFFF : -1;
            }
        }
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
        atomicOr(&g.cand0[(pz - (k / 3 - 1)) * g.DY + (py - (k % 3 - 1))], (1u << k) | (1u << (9 + ((px >> 5) & 15))));
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

constexpr int BFS_THREADS = 1024;      // one block per SM
constexpr int BFS_ROWS_PER_THREAD = 2; // candidate words a thread fetches per batch (issued together)
constexpr int BFS_ITEM_CAP = 5120;     // (row, word) work items a block holds at a time

// Grid-wide barrier for a cooperative (co-resident) launch with one block per SM: one arrival counter that only
// grows (target = barrier number * blocks).  Whether anything happened at a level is NOT carried in this word: a
// block still spinning at barrier L may read the counter after faster blocks arrived at barrier L + 1, so anything
// packed next to the arrivals would mix levels.  bfs_levels_kernel keeps per-level flags instead (news[level & 3]:
// raised before the arrival at `level`, read after it, cleared two levels ahead by block 0).
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ull);
        unsigned long long v;
        do {
            v = *((volatile unsigned long long*)counter);
        } while ((unsigned int)v < target);
        __threadfence();
    }
    __syncthreads();
}

// All levels in one cooperative launch, one block per SM.
//
// cand[p][row] (p = parity of the level that wrote it): bits 0-8 = which of the nine (y,z) neighbour rows of
// `row` received frontier cells, bits 9-24 = in which 32-bit words (word index mod 16).  A (row, word) that
// gets new cells at level L publishes itself to its nine neighbour rows with one atomicOr each; the row's
// scanner at level L+1 reads the word and clears it for level L+3.
//
// Level L+1: rows are dealt to the blocks in groups of eight (one 32-byte sector of candidate words; every
// block sees a thin slice of the whole grid, which balances the wavefront across SMs).  A block fetches its
// candidate words, turns each candidate row into (row, word) items for the words that can change (published
// words dilated by one), compacts them into a shared-memory list (warp scan + one shared atomic per warp) and
// then expands the list ONE LANE PER ITEM: OR of the active neighbour rows' frontier words, x-dilation by
// shifts with the edge bits of the neighbouring words, `& ~blocked`, distance stores for the new bits, and
// the nine publishes.  A wavefront face perpendicular to x touches one word per row, so this costs one lane
// per row instead of a half-warp.
//
// Frontier words are only current for the words a row published; a stale word holds cells discovered two
// levels earlier, whose neighbours are all discovered already, so whatever it contributes is removed by
// `& ~blocked`.
// One BFS level over this block's share of the rows (see bfs_levels_kernel for the data flow).  Returns whether this
// THREAD discovered cells.
__device__ __forceinline__ bool bfs_level_body(const BfsGrid& g, uint32_t level, uint32_t* s_item, uint32_t* s_info,
                                               int* s_warp_sum, int* s_total_p)
{
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int groups = (g.rows + 7) / 8;                         // groups of 8 consecutive rows
    const int slots = (groups + (int)gridDim.x - 1) / (int)gridDim.x * 8;   // row slots of this block
    const uint32_t wmask_all = g.W >= 16 ? 0xFFFFu : ((1u << g.W) - 1u);
    int& s_total = *s_total_p;
    const bool odd = level & 1;
    const uint32_t* __restrict__ fcur = odd ? g.front0 : g.front1;
    uint32_t* __restrict__ fnext = odd ? g.front1 : g.front0;
    uint32_t* __restrict__ cand_in = odd ? g.cand0 : g.cand1;
    uint32_t* __restrict__ cand_out = odd ? g.cand1 : g.cand0;
    bool any_new = false;

    for (int t0 = 0; t0 < slots; t0 += BFS_THREADS * BFS_ROWS_PER_THREAD) {
        // ---- fetch this thread's candidate words (independent loads, one latency) ----
        int rows[BFS_ROWS_PER_THREAD];
        uint32_t words[BFS_ROWS_PER_THREAD];
#pragma unroll
        for (int u = 0; u < BFS_ROWS_PER_THREAD; ++u) {
            const int t = t0 + u * BFS_THREADS + threadIdx.x;
            const int gi = blockIdx.x + (t >> 3) * gridDim.x;
            rows[u] = gi * 8 + (t & 7);
            words[u] = 0;
            if (t < slots && gi < groups && rows[u] < g.rows) {
                words[u] = __ldcg(&cand_in[rows[u]]);
            }
        }
        // ---- items per candidate row: published words dilated by one ----
        uint32_t dmask[BFS_ROWS_PER_THREAD];
        int mine = 0;
#pragma unroll
        for (int u = 0; u < BFS_ROWS_PER_THREAD; ++u) {
            dmask[u] = 0;
            if (words[u] != 0) {
                cand_in[rows[u]] = 0;   // consumed; this array is written again two levels from now
                const int z = rows[u] / g.DY, y = rows[u] - z * g.DY;
                if (!(z == 0 || z == g.DZ - 1 || y == 0 || y == g.DY - 1)) {   // border shell rows are all wall
                    const uint32_t wm = (words[u] >> 9) & 0xFFFFu;
                    // word index is kept mod 16: with more than 16 words per row bit 15 neighbours bit 0
                    uint32_t d = wm | (wm << 1) | (wm >> 1);
                    if (g.W > 16) {
                        d |= (wm >> 15) | (wm << 15);
                    }
                    dmask[u] = d & wmask_all;
                }
            }
            // with W > 16 every word congruent to a set bit becomes an item
            int n = __popc(dmask[u]);
            if (g.W > 16) {
                n = 0;
                for (int w = 0; w < g.W; ++w) n += (dmask[u] >> (w & 15)) & 1u;
            }
            mine += n;
        }
        // ---- block-wide exclusive scan of `mine` ----
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) {
            s_warp_sum[warp] = incl;
        }
        __syncthreads();
        if (warp == 0) {
            int v = s_warp_sum[lane];
            int inc2 = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, inc2, o);
                if (lane >= o) inc2 += x;
            }
            s_warp_sum[lane] = inc2 - v;
            if (lane == 31) {
                s_total = inc2;
            }
        }
        __syncthreads();
        const int my_off = s_warp_sum[warp] + incl - mine;
        const int total = s_total;

        // ---- rounds of at most BFS_ITEM_CAP items ----
        for (int r0 = 0; r0 < total; r0 += BFS_ITEM_CAP) {
            int pos = my_off;
#pragma unroll
            for (int u = 0; u < BFS_ROWS_PER_THREAD; ++u) {
                if (dmask[u] == 0) {
                    continue;
                }
                const uint32_t wm = (words[u] >> 9) & 0xFFFFu;
                const uint32_t nb = words[u] & 0x1FFu;
                for (int w = 0; w < g.W; ++w) {
                    if (!((dmask[u] >> (w & 15)) & 1u)) {
                        continue;
                    }
                    if (pos >= r0 && pos < r0 + BFS_ITEM_CAP) {
                        const uint32_t hl = (w > 0 && ((wm >> ((w - 1) & 15)) & 1u)) ? 1u : 0u;
                        const uint32_t hr = (w + 1 < g.W && ((wm >> ((w + 1) & 15)) & 1u)) ? 1u : 0u;
                        const uint32_t hs = (wm >> (w & 15)) & 1u;
                        s_item[pos - r0] = (uint32_t)rows[u];
                        s_info[pos - r0] = (uint32_t)w | (nb << 8) | (hl << 17) | (hr << 18) | (hs << 19);
                    }
                    ++pos;
                }
            }
            __syncthreads();
            const int n_items = min(BFS_ITEM_CAP, total - r0);
            for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
                const int row = (int)s_item[i];
                const uint32_t info = s_info[i];
                const int w = info & 0xFFu;
                const uint32_t nb = (info >> 8) & 0x1FFu;
                const bool hl = (info >> 17) & 1u, hr = (info >> 18) & 1u, hs = (info >> 19) & 1u;
                const int z = row / g.DY, y = row - z * g.DY;
                const size_t idx = (size_t)row * g.W + w;
                const uint32_t blk = __ldcg(&g.blocked[idx]);
                uint32_t m = 0, ml = 0, mr = 0;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (nb & (1u << k)) {
                        const uint32_t* fr = fcur + (size_t)((z + k / 3 - 1) * g.DY + (y + k % 3 - 1)) * g.W + w;
                        if (hs) m |= __ldcg(fr);
                        if (hl) ml |= __ldcg(fr - 1);
                        if (hr) mr |= __ldcg(fr + 1);
                    }
                }
                const uint32_t dil = m | (m << 1) | (m >> 1) | (ml >> 31) | (mr << 31);
                uint32_t fresh = dil & ~blk;
                fnext[idx] = fresh;
                if (fresh) {
                    g.blocked[idx] = blk | fresh;
                    any_new = true;
                    const uint32_t pub = 1u << (9 + (w & 15));
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        // this row is neighbour k of row (z - dz, y - dy)
                        atomicOr(&cand_out[(z - (k / 3 - 1)) * g.DY + (y - (k % 3 - 1))], pub | (1u << k));
                    }
                    int* d = g.dist + (size_t)row * g.DX + (size_t)w * 32;
                    while (fresh) {
                        const int b = __ffs(fresh) - 1;
                        fresh &= fresh - 1;
                        d[b] = (int)level;
                    }
                }
            }
            __syncthreads();   // the list is rebuilt by the next round / batch
        }
    }
    return any_new;
}

__global__ void __launch_bounds__(BFS_THREADS, 1)
bfs_levels_kernel(const __grid_constant__ BfsGrid g, int max_levels)
{
    __shared__ uint32_t s_item[BFS_ITEM_CAP];   // row
    __shared__ uint32_t s_info[BFS_ITEM_CAP];   // word | nbrmask << 8 | has_left << 17 | has_right << 18
    __shared__ int s_warp_sum[BFS_THREADS / 32];
    __shared__ int s_total;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(&g.ctrl[4]);
    int* news = &g.ctrl[8];   // [4] per-level flags, see grid_barrier

    uint32_t level = 1;
    for (; level <= (uint32_t)max_levels; ++level) {
        const bool any_new = bfs_level_body(g, level, s_item, s_info, s_warp_sum, &s_total);
        const bool block_new = __syncthreads_or(any_new ? 1 : 0) != 0;
        if (threadIdx.x == 0) {
            if (block_new) {
                atomicExch(&news[level & 3], 1);
            }
            if (blockIdx.x == 0) {
                // slot of level + 2: its last readers (level - 2) all arrived at barrier level - 1, which this block
                // has left; its next writers (level + 2) start after barrier level + 1, which needs this block
                atomicExch(&news[(level + 2) & 3], 0);
            }
        }
        grid_barrier(bar, level * gridDim.x);
        if (__ldcg(&news[level & 3]) == 0) {
            break;   // no block discovered anything at this level (every block reads the same flag)
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        g.ctrl[0] = (int)level;
    }
}

// The same wavefront as ONE LAUNCH PER LEVEL, for the planner's banks when they are run behind the caller's back
// (smplgpu_bfs_bank_run_slots_async) while other contexts' expansion rounds share the GPU.  A cooperative launch
// starts only when ALL its blocks fit at once; launches of one level need no co-residency, interleave with anything,
// and cost ~2 us each against tens of milliseconds per bank run.  (Opt-in, SMPLGPU_BANK_STEPWISE=1 with the level
// kernel forced: the banks run on the tile kernel, whose asynchronous form is stepwise by default.)
// ctrl[12] = sticky "finished" flag (set by the first level that finds the previous one empty), mirrored into
// *done_host (page-locked) so that the host can poll without a synchronisation; ctrl[0] = levels run.
__global__ void __launch_bounds__(BFS_THREADS, 1)
bfs_level_step_kernel(const __grid_constant__ BfsGrid g, uint32_t level, volatile int* done_host)
{
    __shared__ uint32_t s_item[BFS_ITEM_CAP];
    __shared__ uint32_t s_info[BFS_ITEM_CAP];
    __shared__ int s_warp_sum[BFS_THREADS / 32];
    __shared__ int s_total;
    int* news = &g.ctrl[8];
    if (__ldcg(&g.ctrl[12]) != 0) {
        return;   // finished at an earlier level
    }
    if (level > 1 && __ldcg(&news[(level - 1) & 3]) == 0) {
        // the previous level discovered nothing: done (every block of this launch sees the same flag)
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            g.ctrl[0] = (int)level - 1;
            g.ctrl[12] = 1;
            if (done_host != nullptr) {
                *done_host = 1;
            }
        }
        return;
    }
    const bool any_new = bfs_level_body(g, level, s_item, s_info, s_warp_sum, &s_total);
    const bool block_new = __syncthreads_or(any_new ? 1 : 0) != 0;
    if (threadIdx.x == 0) {
        if (block_new) {
            atomicExch(&news[level & 3], 1);
        }
        if (blockIdx.x == 0) {
            atomicExch(&news[(level + 2) & 3], 0);   // stream order: level + 2 is launched after this launch ends
        }
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
