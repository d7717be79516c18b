// Kernel (4), temporally blocked: BFS_3D as TILE_K wavefront levels per grid barrier.
//
// bfs_levels_kernel (bfs.cuh) pays one grid barrier plus a chain of dependent L2 round trips per level
// (about 7 us on B200 whatever the amount of work), so a 400^3 grid with 476 levels cannot finish in under
// 3 ms.  Here a level costs one __syncthreads: the grid is cut into tiles of 32 x 16 x 16 cells (one 32-bit
// word of 16 x 16 bit-rows); a block copies a tile plus a halo of TILE_K rows in y and z and one word in x
// (32 x 32 rows of 3 words) into shared memory and runs TILE_K levels on it, one thread per row:
//     frontier' = dilate26(frontier) & ~blocked;   blocked |= frontier'
// Row r of the extended tile is exact for as many levels as it is rows away from the tile's rim (a cell's
// level-s value depends on cells at most s away), so the 16 x 16 interior rows are exact for all TILE_K levels;
// only they are written back (new blocked bits, the frontier after the last level, and the distances, level
// by level).  The result is the level-synchronous BFS, bit for bit.
//
// Between super-steps (TILE_K levels) the tiles talk through global memory:
//   * `front[p]`: the frontier at the start of super-step n (p = n & 1), written by the previous super-step;
//   * `blocked[v]`: two copies; a tile's interior lives in copy ver[tile], a tile that changes writes the
//     other copy and flips its bit, stamped with the super-step, so neighbours reading its rows as halo during
//     the same super-step still take the state at the START of the super-step;
//   * `queue[n % 3]`: the tiles that must run in super-step n.  A tile that ends a super-step with frontier cells
//     queues (once, `flag`) every tile whose interior is within TILE_K cells of one of them, itself included;
//     blocks take queue entries by position, so every block gets the same number of tiles.
// A frontier word that is not rewritten keeps cells discovered earlier; all their neighbours are discovered
// by then, so whatever they contribute is removed by `& ~blocked` (same argument as in bfs.cuh).
// One grid barrier per super-step; the search ends when a queue comes up empty.
//
// Measured on B200 (tools/bfs_only.py, tools/bank_bfs_time.py): single grids run faster than with
// bfs_levels_kernel (64^3: 0.19 vs 0.35 ms, 150^3: 0.44 vs 0.99 ms, 400^3: 3.7 vs 4.3 ms) because they are
// bound by the per-level latency; the stacked planner banks (many wavefronts at once: 0.18 vs 0.15 ms per
// query) are throughput bound and the tile kernel's halo recomputation costs more than the barriers it saves.
// smplgpu.cu picks the kernel accordingly unless smplgpu_bfs_set_mode forces one.
#pragma once

#include "bfs.cuh"

namespace smplgpu {

constexpr int TILE_K = 8;            // levels per super-step = halo width in rows
constexpr int TILE_Y = 16;           // interior rows per tile in y and in z
constexpr int TILE_E = TILE_Y + 2 * TILE_K;   // extended rows per side (32)
constexpr int TILE_THREADS = TILE_E * TILE_E; // one thread per extended row (1024)

struct BfsTiles
{
    int ntx, nty, ntz, ntiles;   // tiles per axis (x in words)
    uint32_t* blocked1;          // second copy of the blocked bitmap
    uint32_t* ver;               // [ntiles] (super-step of the last flip + 1) << 1 | copy holding the tile's interior
    uint32_t* flag;              // [3][ntiles] tile is queued for super-step n (index n % 3)
    int* queue;                  // [3][ntiles] tiles to run in super-step n (index n % 3)
    int* qn;                     // [3] queue lengths, [3] pull cursors
};

// which blocked copy held `tile`'s interior at the START of super-step n (a tile that flips during n stamps n)
__device__ __forceinline__ uint32_t tile_copy(const BfsTiles& t, int tile, int n)
{
    const uint32_t v = __ldcg(&t.ver[tile]);
    return ((v >> 1) == (uint32_t)(n + 1)) ? ((v & 1u) ^ 1u) : (v & 1u);
}

// queue `tile` for the super-step whose queue index is qi (once)
__device__ __forceinline__ void tile_enqueue(const BfsTiles& t, int tile, int qi)
{
    if (atomicExch(&t.flag[(size_t)qi * t.ntiles + tile], 1u) == 0u) {
        t.queue[(size_t)qi * t.ntiles + atomicAdd(&t.qn[qi], 1)] = tile;
    }
}

// seeds also enter the second blocked copy and raise the flags of the tiles around them
__global__ void bfs_tiles_seed_kernel(BfsGrid g, BfsTiles t, const int* __restrict__ seeds, int n_seeds)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seeds) {
        return;
    }
    const int x = seeds[3 * i], y = seeds[3 * i + 1], z = seeds[3 * i + 2];
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) {
        return;
    }
    const int px = x + 1, py = y + 1, pz = z + 1;
    const size_t word = (size_t)(pz * g.DY + py) * g.W + (px >> 5);
    atomicOr(&t.blocked1[word], 1u << (px & 31));
    const int tx = px >> 5, ty = py / TILE_Y, tz = pz / TILE_Y;
    for (int dz = -1; dz <= 1; ++dz) {
        for (int dy = -1; dy <= 1; ++dy) {
            for (int dx = -1; dx <= 1; ++dx) {
                const int ax = tx + dx, ay = ty + dy, az = tz + dz;
                if (ax >= 0 && ay >= 0 && az >= 0 && ax < t.ntx && ay < t.nty && az < t.ntz) {
                    tile_enqueue(t, (az * t.nty + ay) * t.ntx + ax, 0);   // the first super-step
                }
            }
        }
    }
}

// second blocked copy = walls (run after bfs_reset_kernel, same slot mask)
__global__ void bfs_tiles_reset_kernel(BfsGrid g, BfsTiles t, const uint8_t* __restrict__ slot_mask, int slot_dz)
{
    const int total = g.rows * g.W;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const bool on = slot_mask == nullptr || slot_mask[(idx / g.W / g.DY) / slot_dz] != 0;
        if (on) {
            t.blocked1[idx] = g.wall[idx];
        } else {
            // a slot that is not re-run keeps its distances; clear its frontiers so that a tile shared with a
            // re-run slot finds nothing to expand there
            g.front0[idx] = 0;
            g.front1[idx] = 0;
        }
    }
}

#ifndef TILE_BLOCKS_PER_SM
#define TILE_BLOCKS_PER_SM 1
#endif
__global__ void __launch_bounds__(TILE_THREADS, TILE_BLOCKS_PER_SM)
bfs_tiles_kernel(const __grid_constant__ BfsGrid g, const __grid_constant__ BfsTiles t, int max_supersteps)
{
    __shared__ uint32_t sF[2][TILE_THREADS * 3];
    __shared__ uint8_t sAny[2][TILE_THREADS];        // which words of the row have frontier bits (per buffer)
    __shared__ uint8_t sZ[2][TILE_E];                // z-row (= warp) has frontier bits (per buffer)
    __shared__ unsigned int s_act;                   // which of the 27 neighbour directions get activated
    __shared__ int s_next[2];                        // next queue position of this block, double-buffered by tile
                                                     // count: thread 0 writes slot k & 1 for tile k before that tile's
                                                     // first barrier, everyone reads it after one; slot k & 1 is written
                                                     // again for tile k + 2, i.e. after tile k + 1's barriers

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int ry = tid & (TILE_E - 1), rz = tid >> 5;   // TILE_E == 32: a warp is one z-row of the extended tile
    const bool interior_row = ry >= TILE_K && ry < TILE_K + TILE_Y && rz >= TILE_K && rz < TILE_K + TILE_Y;
    // rows away from the interior (0 inside); a halo row j rows out can reach the interior by level TILE_K only
    // through levels <= TILE_K - j, so level s needs just the rows with j <= TILE_K - s
    const int jy = ry < TILE_K ? TILE_K - ry : (ry >= TILE_K + TILE_Y ? ry - (TILE_K + TILE_Y - 1) : 0);
    const int jz = rz < TILE_K ? TILE_K - rz : (rz >= TILE_K + TILE_Y ? rz - (TILE_K + TILE_Y - 1) : 0);
    const int row_out = max(jy, jz);
    const int xwords = (g.DX + 31) / 32;                // words that hold cells
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(&g.ctrl[4]);
    int max_level = 0;

    int n = 0;
    for (; n < max_supersteps; ++n) {
        const int p = n & 1;
        const uint32_t* __restrict__ fcur = p ? g.front1 : g.front0;
        uint32_t* __restrict__ fnext = p ? g.front0 : g.front1;
        const int qi = n % 3, qi_next = (n + 1) % 3, qi_free = (n + 2) % 3;
        const int level0 = n * TILE_K;
        // the active tiles of this super-step, dealt to the blocks by queue position (balanced counts)
        const int q_len = __ldcg(&t.qn[qi]);
#ifdef SMPLGPU_BFS_STATS
        if (blockIdx.x == 0 && tid == 0) atomicAdd(&g.ctrl[3], q_len);
#endif
        if (q_len == 0) {
            break;   // no tile has a frontier left (every block reads the same length)
        }
        if (blockIdx.x == 0 && tid == 0) {
            t.qn[qi_free] = 0;   // its readers ran before the last barrier, its writers start after the next one
        }
        if (blockIdx.x == 0 && tid == 0) {
            t.qn[3 + qi_free] = 0;   // pull cursor of that queue
        }
        {
            // the first tile by block id, the following ones pulled from a shared cursor as blocks become free
            // (tiles differ a lot in cost: a face across x keeps all 1024 rows busy, most others a few)
            int k = 0;   // tiles this block has taken in this super-step
            for (int a = blockIdx.x; a < q_len; ++k) {
                const int tile = __ldcg(&t.queue[(size_t)qi * t.ntiles + a]);
                if (tid == 0) {
                    t.flag[(size_t)qi * t.ntiles + tile] = 0;   // consumed
                    s_next[k & 1] = (int)gridDim.x + atomicAdd(&t.qn[3 + qi], 1);   // next queue position, fetched while this tile runs
                }
#ifdef SMPLGPU_BFS_STATS
                const long long c0 = clock64();
#endif
                const int tx = tile % t.ntx, ty = (tile / t.ntx) % t.nty, tz = tile / (t.ntx * t.nty);
                const int gy = ty * TILE_Y - TILE_K + ry, gz = tz * TILE_Y - TILE_K + rz;
                const bool row_in = gy >= 0 && gy < g.DY && gz >= 0 && gz < g.DZ;
                const size_t grow = (size_t)(row_in ? gz * g.DY + gy : 0);

                // ---- load the frontier of this thread's row (3 words); a tile with no frontier cell in reach
                //      has nothing to do (flags are raised conservatively) ----
                unsigned wm = 0;   // which of the row's three words hold frontier cells
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const int gw = tx - 1 + w;
                    uint32_t f = 0;
                    if (row_in && gw >= 0 && gw < xwords) {
                        f = __ldcg(&fcur[grow * g.W + gw]);
                    }
                    sF[0][tid * 3 + w] = f;
                    sF[1][tid * 3 + w] = 0;
                    wm |= (f != 0 ? 1u : 0u) << w;
                }
                const bool any = wm != 0;
                sAny[0][tid] = (uint8_t)wm;
                sAny[1][tid] = 0;
                {
                    const bool zany = __any_sync(0xffffffffu, any);
                    if (lane == 0) {
                        sZ[0][rz] = zany ? 1 : 0;
                        sZ[1][rz] = 0;
                    }
                }
                if (tid == 0) {
                    s_act = 0;
                }
                const int idle = !__syncthreads_or(any ? 1 : 0);
                a = s_next[k & 1];   // (block-uniform) written before the barrier above
                if (idle) {
                    // nothing else was written to shared memory that the next tile's loads could overtake: sF / sAny /
                    // sZ are rewritten by the thread that wrote them, s_act by thread 0 alone
                    continue;
                }
#ifdef SMPLGPU_BFS_STATS
                if (tid == 0) atomicAdd(&g.ctrl[6], 1);
#endif
                // ---- blocked words (registers), from the copy each owner tile committed last ----
                uint32_t blk[3];
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const int gw = tx - 1 + w;
                    uint32_t bb = 0xFFFFFFFFu;
                    if (row_in && gw >= 0 && gw < xwords) {
                        const int owner = ((gz / TILE_Y) * t.nty + gy / TILE_Y) * t.ntx + gw;
                        const uint32_t* src = tile_copy(t, owner, n) ? t.blocked1 : g.blocked;
                        bb = __ldcg(&src[grow * g.W + gw]);
                    }
                    blk[w] = bb;
                }

#ifdef SMPLGPU_BFS_STATS
                __syncthreads();
                const long long c1 = clock64();
#endif
                // ---- TILE_K levels in shared memory ----
                bool changed = false;      // this thread's interior word gained cells
                uint32_t last = 0;         // interior frontier word after the last level run
                int cur = 0, s = 1;
                for (; s <= TILE_K; ++s) {
                    uint32_t fresh[3] = { 0, 0, 0 };
                    bool got = false;
                    // a warp is one z-row of the tile: nothing to do unless this or an adjacent z-row has frontier
                    const bool zact = jz <= TILE_K - s && (sZ[cur][rz - 1] | sZ[cur][rz] | sZ[cur][rz + 1]);
                    const bool zstale = sZ[cur ^ 1][rz] != 0;   // the buffer written now held cells two levels ago
                    if (zact && row_out <= TILE_K - s) {
                        // any frontier in the 3 x 3 rows around this one?
                        const uint8_t* A = sAny[cur];
                        // words of the 3 x 3 rows around this one that hold frontier cells (a wavefront face across x
                        // touches one word per row: only that word is gathered)
                        const unsigned near = A[tid - 33] | A[tid - 32] | A[tid - 31] | A[tid - 1] | A[tid] | A[tid + 1] |
                                              A[tid + 31] | A[tid + 32] | A[tid + 33];
                        if (near) {
                            const uint32_t* F = sF[cur];
                            uint32_t m[3];
#pragma unroll
                            for (int w = 0; w < 3; ++w) {
                                m[w] = 0;
                                if ((near >> w) & 1u) {
                                    m[w] = F[(tid - 33) * 3 + w] | F[(tid - 32) * 3 + w] | F[(tid - 31) * 3 + w] |
                                           F[(tid - 1) * 3 + w] | F[tid * 3 + w] | F[(tid + 1) * 3 + w] |
                                           F[(tid + 31) * 3 + w] | F[(tid + 32) * 3 + w] | F[(tid + 33) * 3 + w];
                                }
                            }
                            const uint32_t d0 = m[0] | (m[0] << 1) | (m[0] >> 1) | (m[1] << 31);
                            const uint32_t d1 = m[1] | (m[1] << 1) | (m[1] >> 1) | (m[0] >> 31) | (m[2] << 31);
                            const uint32_t d2 = m[2] | (m[2] << 1) | (m[2] >> 1) | (m[1] >> 31);
                            fresh[0] = d0 & ~blk[0];
                            fresh[1] = d1 & ~blk[1];
                            fresh[2] = d2 & ~blk[2];
                            blk[0] |= fresh[0];
                            blk[1] |= fresh[1];
                            blk[2] |= fresh[2];
                            got = (fresh[0] | fresh[1] | fresh[2]) != 0;
                        }
                    }
                    // rows that had cells in the target buffer two levels ago must be cleared
                    if (got || (zstale && sAny[cur ^ 1][tid])) {
                        sF[cur ^ 1][tid * 3] = fresh[0];
                        sF[cur ^ 1][tid * 3 + 1] = fresh[1];
                        sF[cur ^ 1][tid * 3 + 2] = fresh[2];
                        sAny[cur ^ 1][tid] = (uint8_t)((fresh[0] != 0 ? 1u : 0u) | (fresh[1] != 0 ? 2u : 0u) | (fresh[2] != 0 ? 4u : 0u));
                    }
                    const bool zgot = __any_sync(0xffffffffu, got);
                    if (lane == 0 && (zgot || zstale)) {
                        sZ[cur ^ 1][rz] = zgot ? 1 : 0;
                    }
                    if (interior_row) {
                        last = fresh[1];
                        if (fresh[1]) {
                            changed = true;
                            max_level = max(max_level, level0 + s);
                        }
                    }
                    const int alive = __syncthreads_or(got ? 1 : 0);
                    cur ^= 1;
#ifdef SMPLGPU_BFS_STATS
                    if (tid == 0) atomicAdd(&g.ctrl[7], 1);
#endif
                    // distances of the interior word: a row with a few new cells (a wavefront face across x) stores
                    // them itself; dense rows (faces along x) go out warp-wide, lanes = bits, one 128-byte store per row
                    if (zgot && rz >= TILE_K && rz < TILE_K + TILE_Y) {
                        const bool mine = interior_row && fresh[1] != 0;
                        const bool dense = mine && __popc(fresh[1]) > 4;
                        if (mine && !dense) {
                            int* d = g.dist + ((size_t)gz * g.DY + gy) * g.DX + (size_t)tx * 32;
                            uint32_t f = fresh[1];
                            while (f) {
                                d[__ffs(f) - 1] = level0 + s;
                                f &= f - 1;
                            }
                        }
                        uint32_t todo = __ballot_sync(0xffffffffu, dense);
                        while (todo) {
                            const int r = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const uint32_t wk = __shfl_sync(0xffffffffu, fresh[1], r);
                            const int y2 = ty * TILE_Y - TILE_K + r;
                            if ((wk >> lane) & 1u) {
                                g.dist[((size_t)gz * g.DY + y2) * g.DX + (size_t)tx * 32 + lane] = level0 + s;
                            }
                        }
                    }
                    if (!alive) {
                        last = 0;   // nothing was found at level s: the frontier is empty from here on
                        break;
                    }
                }

#ifdef SMPLGPU_BFS_STATS
                const long long c2 = clock64();
#endif
                // ---- write back the interior: blocked into the other copy, frontier for the next super-step ----
                if (__syncthreads_or(changed ? 1 : 0)) {
                    const uint32_t v = tile_copy(t, tile, n);
                    uint32_t* dst = v ? g.blocked : t.blocked1;
                    if (interior_row && row_in && tx < xwords) {
                        dst[grow * g.W + tx] = blk[1];
                    }
                    if (tid == 0) {
                        t.ver[tile] = ((uint32_t)(n + 1) << 1) | (v ^ 1u);
                    }
                }
                if (interior_row && row_in && tx < xwords) {
                    fnext[grow * g.W + tx] = last;
                }
                // tiles whose interior is within TILE_K cells of a remaining frontier cell run next super-step
                if (interior_row && last != 0) {
                    const int iy = ry - TILE_K, iz = rz - TILE_K;
                    const unsigned xs = 2u | ((last & 0x000000FFu) ? 1u : 0u) | ((last & 0xFF000000u) ? 4u : 0u);   // bit dx+1
                    const unsigned ys = 2u | (iy < TILE_K ? 1u : 0u) | (iy >= TILE_Y - TILE_K ? 4u : 0u);
                    const unsigned zs = 2u | (iz < TILE_K ? 1u : 0u) | (iz >= TILE_Y - TILE_K ? 4u : 0u);
                    unsigned act = 0;
#pragma unroll
                    for (int c = 0; c < 27; ++c) {
                        if (((xs >> (c % 3)) & 1u) && ((ys >> ((c / 3) % 3)) & 1u) && ((zs >> (c / 9)) & 1u)) {
                            act |= 1u << c;
                        }
                    }
                    atomicOr(&s_act, act);
                }
                __syncthreads();
                const unsigned act = s_act;
                if (tid < 27 && ((act >> tid) & 1u)) {
                    const int ax = tx + tid % 3 - 1, ay = ty + (tid / 3) % 3 - 1, az = tz + tid / 9 - 1;
                    if (ax >= 0 && ay >= 0 && az >= 0 && ax < t.ntx && ay < t.nty && az < t.ntz) {
                        tile_enqueue(t, (az * t.nty + ay) * t.ntx + ax, qi_next);
                    }
                }
                __syncthreads();   // shared buffers are reused by the next tile
#ifdef SMPLGPU_BFS_STATS
                if (tid == 0) {
                    const long long c3 = clock64();
                    atomicAdd(reinterpret_cast<unsigned long long*>(t.qn + 8), (unsigned long long)(c1 - c0));
                    atomicAdd(reinterpret_cast<unsigned long long*>(t.qn + 10), (unsigned long long)(c2 - c1));
                    atomicAdd(reinterpret_cast<unsigned long long*>(t.qn + 12), (unsigned long long)(c3 - c2));
                }
#endif
            }
        }
#ifdef SMPLGPU_BFS_STATS
        const long long b0 = clock64();
#endif
        grid_barrier(bar, (unsigned int)(n + 1) * gridDim.x);
#ifdef SMPLGPU_BFS_STATS
        if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(t.qn + 14), (unsigned long long)(clock64() - b0));
#endif
    }
    // levels run = deepest level that discovered a cell, + 1 (as bfs_levels_kernel reports it)
    for (int o = 16; o > 0; o >>= 1) {
        max_level = max(max_level, __shfl_xor_sync(0xffffffffu, max_level, o));
    }
    if (lane == 0 && max_level > 0) {
        atomicMax(&g.ctrl[0], max_level + 1);
    }
}

} // namespace smplgpu
