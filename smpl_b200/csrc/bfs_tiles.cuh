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

// A row of the extended tile as ONE 48-bit word: 8 halo cells | the 32 interior cells | 8 halo cells (bit p <-> cell
// x0 - 8 + p).  TILE_K levels can only carry information TILE_K cells far, so 8 cells of the neighbouring bitmap words
// are all a row needs in x -- the first version kept the whole neighbouring words (3 x 32 bits per row, three of
// everything per level).  ncu (profiles/r02b_bfs_metrics.csv) showed that kernel waiting at barriers 54 % of the time:
// per level the whole block waits for the few z-rows that hold the wavefront, each ONE warp walking ~500 dependent
// instructions (27 shared loads, three words of logic); with packed rows a level is 9 shared 64-bit loads and a dozen
// logic operations per active lane.
__device__ __forceinline__ unsigned long long pack_row(uint32_t left, uint32_t mid, uint32_t right)
{
    return (unsigned long long)(left >> 24) | ((unsigned long long)mid << 8) | ((unsigned long long)(right & 0xFFu) << 40);
}

constexpr unsigned long long ROW_MASK = 0xFFFFFFFFFFFFull;   // 48 cells
constexpr int TILE_RPT_LARGE = 4;                             // extended z-rows per warp on large grids

// One block = one tile at a time; a warp owns TILE_RPT z-rows of the extended tile (interleaved, so that the few
// adjacent z-rows a wavefront occupies fall to different warps), lane = y-row.
//   TILE_RPT = 1: 1024 threads, one block per SM.  A tile is finished soonest (every z-row has its own warp): the
//     choice for grids with few tiles per super-step, where the chain of dependent tile steps bounds the run
//     (150^3: 0.43 ms against 0.55-0.61 ms with TILE_RPT = 4).
//   TILE_RPT = 4: 256 threads, four blocks per SM.  With one thread per row a level costs ~1600 cycles however
//     little there is to do -- 32 warps each walk the "nothing here" path and meet at a barrier; here an idle z-row
//     is one flag test, the barrier has 8 participants, one tile's loads, barriers and write-back hide behind the
//     levels of the other three, and the super-step's tiles are dealt in units a quarter as coarse (the grid barrier
//     waits for the slowest block).  The choice for large grids (400^3: 2.74 ms against 3.17 ms).
template <int TILE_RPT>
__global__ void __launch_bounds__(TILE_THREADS / TILE_RPT, TILE_RPT)
bfs_tiles_kernel(const __grid_constant__ BfsGrid g, const __grid_constant__ BfsTiles t, int max_supersteps)
{
    __shared__ unsigned long long sF[2][TILE_THREADS];   // frontier rows, double-buffered by level
    __shared__ uint8_t sZ[2][TILE_E + 2];                // z-row has frontier cells (per buffer); [0] and [33] stay 0
    __shared__ unsigned int s_act;                       // which of the 27 neighbour directions get activated
    __shared__ int s_next[2];                            // next queue position of this block, double-buffered by tile
                                                         // count: thread 0 writes slot k & 1 for tile k before that tile's
                                                         // first barrier, everyone reads it after one; slot k & 1 is written
                                                         // again for tile k + 2, i.e. after tile k + 1's barriers

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int ry = lane;                                 // TILE_E == 32: a warp's lanes are the y-rows of a z-row
    // rows away from the interior (0 inside); a halo row j rows out can reach the interior by level TILE_K only
    // through levels <= TILE_K - j, so level s needs just the rows with j <= TILE_K - s
    const int jy = ry < TILE_K ? TILE_K - ry : (ry >= TILE_K + TILE_Y ? ry - (TILE_K + TILE_Y - 1) : 0);
    const bool interior_y = ry >= TILE_K && ry < TILE_K + TILE_Y;
    int jz[TILE_RPT], row_out[TILE_RPT];
    bool interior_row[TILE_RPT];
#pragma unroll
    for (int r = 0; r < TILE_RPT; ++r) {
        const int rz = warp + (TILE_E / TILE_RPT) * r;
        jz[r] = rz < TILE_K ? TILE_K - rz : (rz >= TILE_K + TILE_Y ? rz - (TILE_K + TILE_Y - 1) : 0);
        row_out[r] = max(jy, jz[r]);
        interior_row[r] = interior_y && rz >= TILE_K && rz < TILE_K + TILE_Y;
    }
    const int xwords = (g.DX + 31) / 32;                // words that hold cells
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(&g.ctrl[4]);
    int max_level = 0;
    if (tid < 2 * (TILE_E + 2)) {
        (&sZ[0][0])[tid] = 0;
    }
    __syncthreads();

    int n = 0;
    for (; n < max_supersteps; ++n) {
        const int p = n & 1;
        const uint32_t* __restrict__ fcur = p ? g.front1 : g.front0;
        uint32_t* __restrict__ fnext = p ? g.front0 : g.front1;
        const int qi = n % 3, qi_next = (n + 1) % 3, qi_free = (n + 2) % 3;
        const int level0 = n * TILE_K;
        // the active tiles of this super-step, dealt to the blocks by queue position
        const int q_len = __ldcg(&t.qn[qi]);
#ifdef SMPLGPU_BFS_STATS
        if (blockIdx.x == 0 && tid == 0) atomicAdd(&g.ctrl[3], q_len);
#endif
        if (q_len == 0) {
            break;   // no tile has a frontier left (every block reads the same length)
        }
        if (blockIdx.x == 0 && tid == 0) {
            t.qn[qi_free] = 0;       // its readers ran before the last barrier, its writers start after the next one
            t.qn[3 + qi_free] = 0;   // pull cursor of that queue
        }
        {
            // the first tile by block id, the following ones pulled from a shared cursor as blocks become free
            // (tiles differ a lot in cost)
            int k = 0;   // tiles this block has taken in this super-step
            for (int a = blockIdx.x; a < q_len; ++k) {
                const int tile = __ldcg(&t.queue[(size_t)qi * t.ntiles + a]);
                if (tid == 0) {
                    t.flag[(size_t)qi * t.ntiles + tile] = 0;   // consumed
                    s_next[k & 1] = (int)gridDim.x + atomicAdd(&t.qn[3 + qi], 1);   // next queue position, fetched while this tile runs
                    s_act = 0;
                }
#ifdef SMPLGPU_BFS_STATS
                const long long c0 = clock64();
#endif
                const int tx = tile % t.ntx, ty = (tile / t.ntx) % t.nty, tz = tile / (t.ntx * t.nty);
                const int gy = ty * TILE_Y - TILE_K + ry;
                bool row_in[TILE_RPT];
                size_t grow[TILE_RPT];
                int gz[TILE_RPT];

                // ---- the frontier of this thread's rows: the interior word and 8 cells of each neighbour ----
                uint32_t fw[TILE_RPT][3];
#pragma unroll
                for (int r = 0; r < TILE_RPT; ++r) {
                    const int rz = warp + (TILE_E / TILE_RPT) * r;
                    gz[r] = tz * TILE_Y - TILE_K + rz;
                    row_in[r] = gy >= 0 && gy < g.DY && gz[r] >= 0 && gz[r] < g.DZ;
                    grow[r] = (size_t)(row_in[r] ? gz[r] * g.DY + gy : 0);
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        const int gw = tx - 1 + w;
                        fw[r][w] = 0;
                        if (row_in[r] && gw >= 0 && gw < xwords) {
                            fw[r][w] = __ldcg(&fcur[grow[r] * g.W + gw]);
                        }
                    }
                }
                bool any0 = false;
#pragma unroll
                for (int r = 0; r < TILE_RPT; ++r) {
                    const int rz = warp + (TILE_E / TILE_RPT) * r;
                    const unsigned long long f0 = pack_row(fw[r][0], fw[r][1], fw[r][2]);
                    sF[0][rz * TILE_E + lane] = f0;
                    sF[1][rz * TILE_E + lane] = 0;
                    const bool zany = __any_sync(0xffffffffu, f0 != 0);
                    if (lane == 0) {
                        sZ[0][rz + 1] = zany ? 1 : 0;
                        sZ[1][rz + 1] = 0;
                    }
                    any0 |= zany;
                }
                const int idle = !__syncthreads_or(any0 ? 1 : 0);
                a = s_next[k & 1];   // (block-uniform) written before the barrier above
                if (idle) {
                    // a tile with no frontier cell in reach has nothing to do (flags are raised conservatively); what
                    // was written to shared memory is rewritten by the same threads for the next tile
                    continue;
                }
#ifdef SMPLGPU_BFS_STATS
                if (tid == 0) atomicAdd(&g.ctrl[6], 1);
#endif
                // ---- blocked cells of the rows (registers), from the copy each owner tile committed last ----
                unsigned long long blk[TILE_RPT];
                {
                    uint32_t bw[TILE_RPT][3];
#pragma unroll
                    for (int r = 0; r < TILE_RPT; ++r) {
#pragma unroll
                        for (int w = 0; w < 3; ++w) {
                            const int gw = tx - 1 + w;
                            uint32_t bb = 0xFFFFFFFFu;
                            if (row_in[r] && gw >= 0 && gw < xwords) {
                                const int owner = ((gz[r] / TILE_Y) * t.nty + gy / TILE_Y) * t.ntx + gw;
                                const uint32_t* src = tile_copy(t, owner, n) ? t.blocked1 : g.blocked;
                                bb = __ldcg(&src[grow[r] * g.W + gw]);
                            }
                            bw[r][w] = bb;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < TILE_RPT; ++r) {
                        blk[r] = pack_row(bw[r][0], bw[r][1], bw[r][2]);
                    }
                }
#ifdef SMPLGPU_BFS_STATS
                __syncthreads();
                const long long c1 = clock64();
#endif
                // ---- TILE_K levels in shared memory ----
                bool changed = false;              // one of this thread's interior words gained cells
                uint32_t last[TILE_RPT] = { };     // interior frontier words after the last level run
                int cur = 0, s = 1;
                for (; s <= TILE_K; ++s) {
                    bool any_fresh = false;
#pragma unroll
                    for (int r = 0; r < TILE_RPT; ++r) {
                        const int rz = warp + (TILE_E / TILE_RPT) * r;
                        // nothing to do unless this or an adjacent z-row has frontier cells
                        const bool zact = jz[r] <= TILE_K - s && (sZ[cur][rz] | sZ[cur][rz + 1] | sZ[cur][rz + 2]);
                        const bool zstale = sZ[cur ^ 1][rz + 1] != 0;   // the buffer written now held cells two levels ago
                        if (!zact && !zstale) {                         // warp-uniform
                            if (interior_row[r]) last[r] = 0;
                            continue;
                        }
                        const int row = rz * TILE_E + lane;
                        unsigned long long fresh = 0;
                        if (zact && row_out[r] <= TILE_K - s) {
                            // the z-rows below / above the extended tile do not exist: rz - 1 < 0 or rz + 1 > 31 only for
                            // rows with jz = TILE_K, which no level computes (jz <= TILE_K - s < TILE_K); the same for y
                            const unsigned long long* F = sF[cur];
                            const unsigned long long m = F[row - 33] | F[row - 32] | F[row - 31] | F[row - 1] | F[row] | F[row + 1] |
                                                         F[row + 31] | F[row + 32] | F[row + 33];
                            fresh = (m | (m << 1) | (m >> 1)) & ~blk[r] & ROW_MASK;
                            blk[r] |= fresh;
                        }
                        const bool zgot = __any_sync(0xffffffffu, fresh != 0);
                        sF[cur ^ 1][row] = fresh;      // the whole z-row is rewritten, so a clear flag means clear rows
                        if (lane == 0) {
                            sZ[cur ^ 1][rz + 1] = zgot ? 1 : 0;
                        }
                        any_fresh |= zgot;
                        const uint32_t fresh_in = (uint32_t)(fresh >> 8);
                        if (interior_row[r]) {
                            last[r] = fresh_in;
                            if (fresh_in) {
                                changed = true;
                                max_level = max(max_level, level0 + s);
                            }
                        }
                        // distances of the interior word: a row with a few new cells (a wavefront face across x) stores
                        // them itself; dense rows (faces along x) go out warp-wide, lanes = bits, one 128-byte store per row
                        if (zgot && rz >= TILE_K && rz < TILE_K + TILE_Y) {
                            const bool mine = interior_row[r] && fresh_in != 0;
                            const bool dense = mine && __popc(fresh_in) > 4;
                            if (mine && !dense) {
                                int* d = g.dist + ((size_t)gz[r] * g.DY + gy) * g.DX + (size_t)tx * 32;
                                uint32_t f = fresh_in;
                                while (f) {
                                    d[__ffs(f) - 1] = level0 + s;
                                    f &= f - 1;
                                }
                            }
                            uint32_t todo = __ballot_sync(0xffffffffu, dense);
                            while (todo) {
                                const int rr = __ffs(todo) - 1;
                                todo &= todo - 1;
                                const uint32_t wk = __shfl_sync(0xffffffffu, fresh_in, rr);
                                const int y2 = ty * TILE_Y - TILE_K + rr;
                                if ((wk >> lane) & 1u) {
                                    g.dist[((size_t)gz[r] * g.DY + y2) * g.DX + (size_t)tx * 32 + lane] = level0 + s;
                                }
                            }
                        }
                    }
                    const int alive = __syncthreads_or(any_fresh ? 1 : 0);
                    cur ^= 1;
#ifdef SMPLGPU_BFS_STATS
                    if (tid == 0) atomicAdd(&g.ctrl[7], 1);
#endif
                    if (!alive) {
#pragma unroll
                        for (int r = 0; r < TILE_RPT; ++r) last[r] = 0;   // nothing was found at level s: the frontier is empty from here on
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
#pragma unroll
                    for (int r = 0; r < TILE_RPT; ++r) {
                        if (interior_row[r] && row_in[r] && tx < xwords) {
                            dst[grow[r] * g.W + tx] = (uint32_t)(blk[r] >> 8);
                        }
                    }
                    if (tid == 0) {
                        t.ver[tile] = ((uint32_t)(n + 1) << 1) | (v ^ 1u);
                    }
                }
                unsigned act = 0;
#pragma unroll
                for (int r = 0; r < TILE_RPT; ++r) {
                    if (interior_row[r] && row_in[r] && tx < xwords) {
                        fnext[grow[r] * g.W + tx] = last[r];
                    }
                    // tiles whose interior is within TILE_K cells of a remaining frontier cell run next super-step
                    if (interior_row[r] && last[r] != 0) {
                        const int rz = warp + (TILE_E / TILE_RPT) * r;
                        const int iy = ry - TILE_K, iz = rz - TILE_K;
                        const unsigned xs = 2u | ((last[r] & 0x000000FFu) ? 1u : 0u) | ((last[r] & 0xFF000000u) ? 4u : 0u);   // bit dx+1
                        const unsigned ys = 2u | (iy < TILE_K ? 1u : 0u) | (iy >= TILE_Y - TILE_K ? 4u : 0u);
                        const unsigned zs = 2u | (iz < TILE_K ? 1u : 0u) | (iz >= TILE_Y - TILE_K ? 4u : 0u);
#pragma unroll
                        for (int c = 0; c < 27; ++c) {
                            if (((xs >> (c % 3)) & 1u) && ((ys >> ((c / 3) % 3)) & 1u) && ((zs >> (c / 9)) & 1u)) {
                                act |= 1u << c;
                            }
                        }
                    }
                }
                if (act != 0) {
                    atomicOr(&s_act, act);
                }
                __syncthreads();
                const unsigned acts = s_act;
                if (tid < 27 && ((acts >> tid) & 1u)) {
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
