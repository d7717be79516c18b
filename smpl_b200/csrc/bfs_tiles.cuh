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
// only they are written back (new blocked bits, the frontier after the last level, and the distances of the
// cells found, each with its level).  The result is the level-synchronous BFS, bit for bit.
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
// Measured on B200 (tools/bfs_only.py, tools/bank_bfs_time.py; end of round 2): single grids run faster than
// with bfs_levels_kernel (64^3: 0.16 vs 0.37 ms, 150^3: 0.29 vs 1.05 ms, 400^3: 2.06 vs 4.56 ms) because they
// are bound by the per-level latency.  On the stacked planner banks the tile kernel is faster in isolation too
// (0.087 vs 0.158 ms per query) but not inside the planner, where its cooperative launch has to find room next
// to the other contexts' expansion rounds, so smplgpu.cu uses it for single grids and
// the level kernel for the banks unless smplgpu_bfs_set_mode forces one.
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
    // Tile-major bitmaps, private to the tile kernels: tile (tx, ty, tz) owns 256 consecutive words, word
    // (z & 15) * 16 + (y & 15) = bitmap word tx of row (y, z).  A warp that fetches the 32 y-rows of an extended tile's
    // z-row reads 8 + 16 + 8 consecutive words of three tiles -- four 32-byte sectors per request where the row-major
    // bitmaps (stride W words between y-rows) cost 32, and a tile plus halo is 768 sectors instead of 6144.  Rows
    // beyond DY / DZ inside the last tiles are padding: blocked = all ones, frontier = 0.
    uint32_t* tb[2];             // the two copies of the blocked bitmap
    uint32_t* tf[2];             // frontier at the start of super-step n: tf[n & 1]
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

constexpr int TILE_NONE = -0x40000000;         // "no such tile" (tile index - 1 can be -1 for a real row)
constexpr int TILE_WORDS = TILE_Y * TILE_Y;   // words of one tile in the tile-major bitmaps

__device__ __forceinline__ size_t tile_word(int tile, int zi, int yi)
{
    return (size_t)tile * TILE_WORDS + zi * TILE_Y + yi;
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
    const int tx = px >> 5, ty = py / TILE_Y, tz = pz / TILE_Y;
    const size_t word = tile_word((tz * t.nty + ty) * t.ntx + tx, pz % TILE_Y, py % TILE_Y);
    atomicOr(&t.tb[0][word], 1u << (px & 31));
    atomicOr(&t.tb[1][word], 1u << (px & 31));
    atomicOr(&t.tf[0][word], 1u << (px & 31));
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

// both blocked copies = walls, both frontiers empty (run after bfs_reset_kernel, same slot mask).  Padding rows and the
// slots of a stacked bank that are not re-run hold "everything blocked": nothing to discover there.
__global__ void bfs_tiles_reset_kernel(BfsGrid g, BfsTiles t, const uint8_t* __restrict__ slot_mask, int slot_dz)
{
    const size_t total = (size_t)t.ntiles * TILE_WORDS;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int tile = (int)(idx / TILE_WORDS), in = (int)(idx % TILE_WORDS);
        const int tx = tile % t.ntx, ty = (tile / t.ntx) % t.nty, tz = tile / (t.ntx * t.nty);
        const int gy = ty * TILE_Y + in % TILE_Y, gz = tz * TILE_Y + in / TILE_Y;
        uint32_t w = 0xFFFFFFFFu;
        if (gy < g.DY && gz < g.DZ && (slot_mask == nullptr || slot_mask[gz / slot_dz] != 0)) {
            w = g.wall[(size_t)(gz * g.DY + gy) * g.W + tx];
        }
        t.tb[0][idx] = w;
        t.tb[1][idx] = w;
        t.tf[0][idx] = 0;
        t.tf[1][idx] = 0;
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
#ifdef SMPLGPU_BFS_STATS
__device__ unsigned long long bfs_dbg[4][2048];   // per super-step: queue length, slowest block's busy cycles, most tiles a block took, sum of busy cycles
#endif
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
bfs_tiles_kernel(const __grid_constant__ BfsGrid g, const __grid_constant__ BfsTiles t, int max_supersteps, int first_step,
                 int stepwise, volatile int* done_host)
{
    __shared__ unsigned long long sF[2][TILE_THREADS];   // frontier rows, double-buffered by level
    __shared__ uint8_t sZ[2][TILE_E + 2];                // z-row has frontier cells (per buffer); [0] and [33] stay 0
    // Distances leave the tile AFTER its levels: cells found this super-step (sNew) and the level of each, minus one, as
    // three bit planes (sLev), per interior row.  Storing them level by level -- per-lane loops for rows with a few new
    // cells, warp-wide stores for dense rows, all between two block barriers -- cost 0.8 of 2.6 ms at 400^3: the slowest
    // warp's stores were on the critical path of every level of every tile.
    __shared__ uint32_t sNew[TILE_WORDS];
    __shared__ uint32_t sLev[3][TILE_WORDS];
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
    int jz[TILE_RPT], row_out[TILE_RPT], irow[TILE_RPT];
    bool interior_row[TILE_RPT];
#pragma unroll
    for (int r = 0; r < TILE_RPT; ++r) {
        const int rz = warp + (TILE_E / TILE_RPT) * r;
        jz[r] = rz < TILE_K ? TILE_K - rz : (rz >= TILE_K + TILE_Y ? rz - (TILE_K + TILE_Y - 1) : 0);
        row_out[r] = max(jy, jz[r]);
        interior_row[r] = interior_y && rz >= TILE_K && rz < TILE_K + TILE_Y;
        irow[r] = (rz - TILE_K) * TILE_Y + (ry - TILE_K);
    }
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(&g.ctrl[4]);
    int max_level = 0;
    if (tid < 2 * (TILE_E + 2)) {
        (&sZ[0][0])[tid] = 0;
    }
    __syncthreads();

    // stepwise: ONE super-step per launch (first_step), no grid barrier and therefore no need for all blocks to be
    // resident at once -- for runs queued behind the caller's back while other contexts' kernels share the GPU (a
    // cooperative launch waits until every block fits at the same moment).  Same planner throughput as the cooperative
    // form (2320-2460 vs 2370-2410 queries/s), ~19 launches per 150^3 bank run.
    // ctrl[12] = sticky "finished" flag, mirrored into *done_host (page-locked) for the host's poll.
    if (stepwise && __ldcg(&g.ctrl[12]) != 0) {
        return;
    }
    int n = first_step;
    for (; n < max_supersteps; ++n) {
        const int p = n & 1;
        const uint32_t* __restrict__ fcur = t.tf[p];
        uint32_t* __restrict__ fnext = t.tf[p ^ 1];
        const int qi = n % 3, qi_next = (n + 1) % 3, qi_free = (n + 2) % 3;
        const int level0 = n * TILE_K;
        // the active tiles of this super-step, dealt to the blocks by queue position
        const int q_len = __ldcg(&t.qn[qi]);
#ifdef SMPLGPU_BFS_STATS
        if (blockIdx.x == 0 && tid == 0) atomicAdd(&g.ctrl[3], q_len);
        const long long ss0 = clock64();
        int taken = 0;
#endif
        if (q_len == 0) {
            if (stepwise && blockIdx.x == 0 && tid == 0) {
                g.ctrl[12] = 1;
                if (done_host != nullptr) {
                    *done_host = 1;
                }
            }
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
                ++taken;
#endif
                const int tx = tile % t.ntx, ty = (tile / t.ntx) % t.nty, tz = tile / (t.ntx * t.nty);
                // the tile that owns this thread's rows (y part) and the rows' place in its 16 x 16 block
                const int oty = ty + (ry < TILE_K ? -1 : (ry >= TILE_K + TILE_Y ? 1 : 0));
                const int yi = (ry + TILE_K) & (TILE_Y - 1);
                int own[TILE_RPT];       // owner tile of the row in x-column tx - 1 (+1, +2 for the other two), or TILE_NONE

                // ---- the frontier of this thread's rows: the interior word and 8 cells of each neighbour ----
                uint32_t fw[TILE_RPT][3];
                unsigned cps = 0;   // bit 3 r + w: which blocked copy the owner of row r, x-column w committed last (the
                                    // version words are fetched with the frontier, one L2 round trip ahead of the blocked rows)
#pragma unroll
                for (int r = 0; r < TILE_RPT; ++r) {
                    const int rz = warp + (TILE_E / TILE_RPT) * r;
                    const int otz = tz + (rz < TILE_K ? -1 : (rz >= TILE_K + TILE_Y ? 1 : 0));
                    const bool in = oty >= 0 && oty < t.nty && otz >= 0 && otz < t.ntz;
                    own[r] = in ? (otz * t.nty + oty) * t.ntx + tx - 1 : TILE_NONE;
                    const int wi = ((rz + TILE_K) & (TILE_Y - 1)) * TILE_Y + yi;
#pragma unroll
                    for (int w = 0; w < 3; ++w) {
                        const int gw = tx - 1 + w;
                        fw[r][w] = 0;
                        if (in && gw >= 0 && gw < t.ntx) {
                            fw[r][w] = __ldcg(&fcur[(size_t)(own[r] + w) * TILE_WORDS + wi]);
                            cps |= tile_copy(t, own[r] + w, n) << (3 * r + w);
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
                    if (interior_row[r]) {
                        sNew[irow[r]] = 0;
                        sLev[0][irow[r]] = 0;
                        sLev[1][irow[r]] = 0;
                        sLev[2][irow[r]] = 0;
                    }
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
                        const int rz = warp + (TILE_E / TILE_RPT) * r;
                        const int wi = ((rz + TILE_K) & (TILE_Y - 1)) * TILE_Y + yi;
#pragma unroll
                        for (int w = 0; w < 3; ++w) {
                            const int gw = tx - 1 + w;
                            uint32_t bb = 0xFFFFFFFFu;
                            if (own[r] != TILE_NONE && gw >= 0 && gw < t.ntx) {
                                bb = __ldcg(&t.tb[(cps >> (3 * r + w)) & 1u][(size_t)(own[r] + w) * TILE_WORDS + wi]);
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
                        if (zact) {   // warp-uniform
                            // dilation in z from shared memory (three rows), in y by shuffles, in x by shifts -- nine
                            // 64-bit shared loads per row kept the shared-memory pipe busy 576 cycles per level of a
                            // fully active tile.  zact implies 0 < rz < 31 (the rim rows have jz = TILE_K, which no
                            // level computes); lanes 0 and 31 get their own value back from the shuffles.
                            const unsigned long long* F = sF[cur];
                            unsigned long long m = F[row - TILE_E] | F[row] | F[row + TILE_E];
                            m |= __shfl_up_sync(0xffffffffu, m, 1) | __shfl_down_sync(0xffffffffu, m, 1);
                            if (row_out[r] <= TILE_K - s) {
                                fresh = (m | (m << 1) | (m >> 1)) & ~blk[r] & ROW_MASK;
                                blk[r] |= fresh;
                            }
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
                                // the level of every new cell, bit-sliced (only this thread touches its rows' words
                                // until the barrier after the last level)
                                changed = true;
                                max_level = max(max_level, level0 + s);
                                sNew[irow[r]] |= fresh_in;
                                if ((s - 1) & 1) sLev[0][irow[r]] |= fresh_in;
                                if ((s - 1) & 2) sLev[1][irow[r]] |= fresh_in;
                                if ((s - 1) & 4) sLev[2][irow[r]] |= fresh_in;
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
                    {
                        // distances: a warp takes 8 * TILE_RPT interior rows; lanes = cells, one store of up to 128 bytes
                        // per row that gained cells
                        constexpr int ROWS_PER_WARP = TILE_WORDS / (TILE_THREADS / TILE_RPT / 32);
                        const int first = warp * ROWS_PER_WARP;
                        const uint32_t mine = lane < ROWS_PER_WARP ? sNew[first + lane] : 0u;
                        uint32_t todo = __ballot_sync(0xffffffffu, mine != 0);
                        while (todo) {
                            const int rr = __ffs(todo) - 1;
                            todo &= todo - 1;
                            const int ir = first + rr;
                            const uint32_t nw = __shfl_sync(0xffffffffu, mine, rr);
                            const uint32_t l0 = sLev[0][ir], l1 = sLev[1][ir], l2 = sLev[2][ir];
                            if ((nw >> lane) & 1u) {
                                const int lev = 1 + (int)(((l0 >> lane) & 1u) | (((l1 >> lane) & 1u) << 1) | (((l2 >> lane) & 1u) << 2));
                                const int y2 = ty * TILE_Y + (ir % TILE_Y), z2 = tz * TILE_Y + ir / TILE_Y;
                                g.dist[((size_t)z2 * g.DY + y2) * g.DX + (size_t)tx * 32 + lane] = level0 + lev;
                            }
                        }
                    }
                    const uint32_t v = tile_copy(t, tile, n);
                    uint32_t* dst = t.tb[v ^ 1u];
#pragma unroll
                    for (int r = 0; r < TILE_RPT; ++r) {
                        if (interior_row[r]) {
                            const int rz = warp + (TILE_E / TILE_RPT) * r;
                            dst[tile_word(tile, rz - TILE_K, ry - TILE_K)] = (uint32_t)(blk[r] >> 8);
                        }
                    }
                    if (tid == 0) {
                        t.ver[tile] = ((uint32_t)(n + 1) << 1) | (v ^ 1u);
                    }
                }
                unsigned act = 0;
#pragma unroll
                for (int r = 0; r < TILE_RPT; ++r) {
                    if (interior_row[r]) {
                        const int rz = warp + (TILE_E / TILE_RPT) * r;
                        fnext[tile_word(tile, rz - TILE_K, ry - TILE_K)] = last[r];
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
        if (tid == 0 && n < 2048) {
            bfs_dbg[0][n] = (unsigned long long)q_len;
            atomicMax(&bfs_dbg[1][n], (unsigned long long)(b0 - ss0));
            atomicMax(&bfs_dbg[2][n], (unsigned long long)taken);
            atomicAdd(&bfs_dbg[3][n], (unsigned long long)(b0 - ss0));
        }
#endif
        if (stepwise) {
            break;   // the next launch is the barrier
        }
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

// ONE WARP PER TILE, the extended tile in registers: lane = y-row, register index = z-row (the z loop is unrolled), a
// level is  m = F[z-1] | F[z] | F[z+1];  m |= shfl_up(m) | shfl_down(m);  fresh = (m | m << 1 | m >> 1) & ~B[z]  -- no
// shared memory and NO block barrier.  ncu on bfs_tiles_kernel<4> (profiles/r02c_bfs_metrics.csv): 18.6 of the stall
// cycles per issued instruction are block barriers -- eight warps meeting once per level although a level is a few
// dozen instructions of work per z-row.  Here the frontier is updated in place (ascending z, the old value of the
// previous z-row carried in a register), an idle z-row costs one vote, twelve tiles are in flight per SM instead of
// four, and the tiles of a super-step are dealt warp by warp.  Protocol between the tiles (queues, flags, the two
// blocked copies and their version stamps): exactly bfs_tiles_kernel's, see the head of this file.
constexpr int WTILE_WARPS = 4;        // warps per block
constexpr int WTILE_BLOCKS_PER_SM = 2;
constexpr int WTILE_GROUP = 8;      // z-rows handled as one straight-line group

__global__ void __launch_bounds__(WTILE_WARPS * 32, WTILE_BLOCKS_PER_SM)
bfs_warp_tiles_kernel(const __grid_constant__ BfsGrid g, const __grid_constant__ BfsTiles t, int max_supersteps)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * WTILE_WARPS + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * WTILE_WARPS;
    const int ry = lane;
    const int jy = ry < TILE_K ? TILE_K - ry : (ry >= TILE_K + TILE_Y ? ry - (TILE_K + TILE_Y - 1) : 0);
    const bool interior_y = ry >= TILE_K && ry < TILE_K + TILE_Y;
    const int oy = ry < TILE_K ? 0 : (ry < TILE_K + TILE_Y ? 1 : 2);   // which of the three tiles in y owns this row
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(&g.ctrl[4]);
    int max_level = 0;

    int n = 0;
    for (; n < max_supersteps; ++n) {
        const int p = n & 1;
        const uint32_t* __restrict__ fcur = t.tf[p];
        uint32_t* __restrict__ fnext = t.tf[p ^ 1];
        const int qi = n % 3, qi_next = (n + 1) % 3, qi_free = (n + 2) % 3;
        const int level0 = n * TILE_K;
        const int q_len = __ldcg(&t.qn[qi]);
        if (q_len == 0) {
            break;
        }
        if (gwarp == 0 && lane == 0) {
            t.qn[qi_free] = 0;
            t.qn[3 + qi_free] = 0;
        }
        for (int a = gwarp; a < q_len;) {
            const int tile = __ldcg(&t.queue[(size_t)qi * t.ntiles + a]);
            {
                int nxt = 0;
                if (lane == 0) {
                    t.flag[(size_t)qi * t.ntiles + tile] = 0;               // consumed
                    nxt = nwarps + atomicAdd(&t.qn[3 + qi], 1);             // next queue position of this warp
                }
                a = __shfl_sync(FULL, nxt, 0);
            }
#ifdef SMPLGPU_BFS_STATS
            const long long c0 = clock64();
            if (lane == 0) atomicAdd(&g.ctrl[3], 1);
#endif
            const int tx = tile % t.ntx, ty = (tile / t.ntx) % t.nty, tz = tile / (t.ntx * t.nty);
            const int gy = ty * TILE_Y - TILE_K + ry;
            const int gz0 = tz * TILE_Y - TILE_K;
            // the tile that owns this lane's rows in y, and the rows' place in its 16 x 16 block
            const int oty = ty + oy - 1;
            const bool y_in = oty >= 0 && oty < t.nty;
            const int yi = (ry + TILE_K) & (TILE_Y - 1);

            // ---- frontier rows ----
            unsigned long long F[TILE_E];
            bool anyf = false;
#pragma unroll
            for (int z = 0; z < TILE_E; ++z) {
                const int otz = tz + (z < TILE_K ? -1 : (z < TILE_K + TILE_Y ? 0 : 1));
                const bool row_in = y_in && otz >= 0 && otz < t.ntz;
                const size_t base = (size_t)(row_in ? (otz * t.nty + oty) * t.ntx + tx : 0) * TILE_WORDS +
                                    ((z + TILE_K) & (TILE_Y - 1)) * TILE_Y + yi;
                uint32_t w0 = 0, w1 = 0, w2 = 0;
                if (row_in) {
                    if (tx >= 1) w0 = __ldcg(&fcur[base - TILE_WORDS]);
                    w1 = __ldcg(&fcur[base]);
                    if (tx + 1 < t.ntx) w2 = __ldcg(&fcur[base + TILE_WORDS]);
                }
                F[z] = pack_row(w0, w1, w2);
                anyf |= F[z] != 0;
            }
            if (!__any_sync(FULL, anyf)) {
                continue;   // no frontier cell in reach (flags are raised conservatively)
            }

            // ---- blocked rows, each from the copy its owner tile committed last (27 owners: one lane each) ----
            unsigned long long B[TILE_E];
            {
                unsigned copy_bit = 0;
                if (lane < 27) {
                    const int ax = tx + lane % 3 - 1, ay = ty + (lane / 3) % 3 - 1, az = tz + lane / 9 - 1;
                    if (ax >= 0 && ay >= 0 && az >= 0 && ax < t.ntx && ay < t.nty && az < t.ntz) {
                        copy_bit = tile_copy(t, (az * t.nty + ay) * t.ntx + ax, n);
                    }
                }
                const unsigned copies = __ballot_sync(FULL, copy_bit != 0);   // bit (oz * 3 + oy) * 3 + ox
#pragma unroll
                for (int z = 0; z < TILE_E; ++z) {
                    const int oz = z < TILE_K ? 0 : (z < TILE_K + TILE_Y ? 1 : 2);
                    const int otz = tz + oz - 1;
                    const bool row_in = y_in && otz >= 0 && otz < t.ntz;
                    const size_t base = (size_t)(row_in ? (otz * t.nty + oty) * t.ntx + tx : 0) * TILE_WORDS +
                                        ((z + TILE_K) & (TILE_Y - 1)) * TILE_Y + yi;
                    const unsigned sel = copies >> ((oz * 3 + oy) * 3);
                    uint32_t w0 = 0xFFFFFFFFu, w1 = 0xFFFFFFFFu, w2 = 0xFFFFFFFFu;
                    if (row_in) {
                        if (tx >= 1) w0 = __ldcg(&t.tb[sel & 1u][base - TILE_WORDS]);
                        w1 = __ldcg(&t.tb[(sel >> 1) & 1u][base]);
                        if (tx + 1 < t.ntx) w2 = __ldcg(&t.tb[(sel >> 2) & 1u][base + TILE_WORDS]);
                    }
                    B[z] = pack_row(w0, w1, w2);
                }
            }

#ifdef SMPLGPU_BFS_STATS
            const long long c1 = clock64();
            if (lane == 0) atomicAdd(&g.ctrl[6], 1);
#endif
            // ---- TILE_K levels in registers ----
            // z-rows in groups of WTILE_GROUP: one vote decides whether a group has anything in reach, and inside an
            // active group everything is straight-line code -- the eight rows' ORs, shuffles and shifts are independent
            // chains a single warp can keep in flight (with a branch per z-row a level cost 32 dependent vote -> shuffle
            // -> logic chains in sequence: 5.8 ms at 400^3 against bfs_tiles_kernel's 2.76 ms)
            bool changed = false;
#pragma unroll 1
            for (int s = 1; s <= TILE_K; ++s) {
                bool any_fresh = false;
                unsigned long long prev = 0;   // the row below the group, as it was before this level
#pragma unroll
                for (int z0 = 0; z0 < TILE_E; z0 += WTILE_GROUP) {
                    unsigned long long reach = prev;
#pragma unroll
                    for (int k = 0; k < WTILE_GROUP; ++k) {
                        reach |= F[z0 + k];
                    }
                    if (z0 + WTILE_GROUP < TILE_E) {
                        reach |= F[z0 + WTILE_GROUP];
                    }
                    if (!__any_sync(FULL, reach != 0)) {   // warp-uniform
                        prev = 0;   // = F[z0 + WTILE_GROUP - 1]; the group's rows are zero and stay zero
                        continue;
                    }
                    unsigned long long fresh[WTILE_GROUP];
                    uint32_t fresh_any_in = 0;
#pragma unroll
                    for (int k = 0; k < WTILE_GROUP; ++k) {
                        const int z = z0 + k;
                        const int jz = z < TILE_K ? TILE_K - z : (z >= TILE_K + TILE_Y ? z - (TILE_K + TILE_Y - 1) : 0);
                        unsigned long long m = (k == 0 ? prev : F[z - 1]) | F[z];
                        if (z + 1 < TILE_E) {
                            m |= F[z + 1];
                        }
                        // lanes 0 and 31 get their own value back from the shuffle: no row beyond the extended tile
                        m |= __shfl_up_sync(FULL, m, 1) | __shfl_down_sync(FULL, m, 1);
                        unsigned long long f = (m | (m << 1) | (m >> 1)) & ~B[z] & ROW_MASK;
                        // a halo row j rows out can reach the interior by level TILE_K only through levels <= TILE_K - j
                        if (max(jy, jz) > TILE_K - s) {
                            f = 0;
                        }
                        fresh[k] = f;
                        if (z >= TILE_K && z < TILE_K + TILE_Y && interior_y) {
                            fresh_any_in |= (uint32_t)(f >> 8);
                        }
                    }
                    prev = F[z0 + WTILE_GROUP - 1];
                    unsigned long long got = 0;
#pragma unroll
                    for (int k = 0; k < WTILE_GROUP; ++k) {
                        B[z0 + k] |= fresh[k];
                        F[z0 + k] = fresh[k];
                        got |= fresh[k];
                    }
                    any_fresh |= got != 0;
                    if (z0 + WTILE_GROUP > TILE_K && z0 < TILE_K + TILE_Y) {
                        // distances of the interior words: sparse rows store their cells themselves, dense rows go out
                        // warp-wide (lanes = bits, one 128-byte store per row)
                        if (fresh_any_in) {
                            changed = true;
                            max_level = max(max_level, level0 + s);
                        }
                        if (__any_sync(FULL, fresh_any_in != 0)) {
#pragma unroll
                            for (int k = 0; k < WTILE_GROUP; ++k) {
                                const int z = z0 + k;
                                if (z >= TILE_K && z < TILE_K + TILE_Y) {
                                    const int gz = gz0 + z;
                                    const uint32_t fresh_in = interior_y ? (uint32_t)(fresh[k] >> 8) : 0u;
                                    const bool dense = __popc(fresh_in) > 4;
                                    if (fresh_in != 0 && !dense) {
                                        int* d = g.dist + ((size_t)gz * g.DY + gy) * g.DX + (size_t)tx * 32;
                                        uint32_t f = fresh_in;
                                        while (f) {
                                            d[__ffs(f) - 1] = level0 + s;
                                            f &= f - 1;
                                        }
                                    }
                                    uint32_t todo = __ballot_sync(FULL, dense);
                                    while (todo) {
                                        const int rr = __ffs(todo) - 1;
                                        todo &= todo - 1;
                                        const uint32_t wk = __shfl_sync(FULL, fresh_in, rr);
                                        const int y2 = ty * TILE_Y - TILE_K + rr;
                                        if ((wk >> lane) & 1u) {
                                            g.dist[((size_t)gz * g.DY + y2) * g.DX + (size_t)tx * 32 + lane] = level0 + s;
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
#ifdef SMPLGPU_BFS_STATS
                if (lane == 0) atomicAdd(&g.ctrl[7], 1);
#endif
                if (!__any_sync(FULL, any_fresh)) {
                    break;   // nothing found at level s: every F[z] is zero, the frontier is empty from here on
                }
            }

#ifdef SMPLGPU_BFS_STATS
            const long long c2 = clock64();
#endif
            // ---- write back the interior: blocked into the other copy, frontier for the next super-step ----
            const bool tile_changed = __any_sync(FULL, changed);
            uint32_t* dst = nullptr;
            if (tile_changed) {
                const uint32_t v = tile_copy(t, tile, n);
                dst = t.tb[v ^ 1u];
                if (lane == 0) {
                    t.ver[tile] = ((uint32_t)(n + 1) << 1) | (v ^ 1u);
                }
            }
            unsigned act = 0;
            if (interior_y) {
#pragma unroll
                for (int z = TILE_K; z < TILE_K + TILE_Y; ++z) {
                    {
                        const size_t idx = tile_word(tile, z - TILE_K, ry - TILE_K);
                        if (tile_changed) {
                            dst[idx] = (uint32_t)(B[z] >> 8);
                        }
                        const uint32_t last = (uint32_t)(F[z] >> 8);
                        fnext[idx] = last;
                        if (last != 0) {
                            // tiles whose interior is within TILE_K cells of a remaining frontier cell run next
                            const int iy = ry - TILE_K, iz = z - TILE_K;
                            const unsigned xs = 2u | ((last & 0x000000FFu) ? 1u : 0u) | ((last & 0xFF000000u) ? 4u : 0u);
                            const unsigned ys = 2u | (iy < TILE_K ? 1u : 0u) | (iy >= TILE_Y - TILE_K ? 4u : 0u);
                            const unsigned zs = 2u | (iz < TILE_K ? 1u : 0u) | (iz >= TILE_Y - TILE_K ? 4u : 0u);
                            // bit (dz * 3 + dy) * 3 + dx: the outer product of the three 3-bit sets
                            const unsigned xy = (ys & 1u ? xs : 0u) | (ys & 2u ? xs << 3 : 0u) | (ys & 4u ? xs << 6 : 0u);
                            act |= (zs & 1u ? xy : 0u) | (zs & 2u ? xy << 9 : 0u) | (zs & 4u ? xy << 18 : 0u);
                        }
                    }
                }
            }
            act = __reduce_or_sync(FULL, act);
            if (lane < 27 && ((act >> lane) & 1u)) {
                const int ax = tx + lane % 3 - 1, ay = ty + (lane / 3) % 3 - 1, az = tz + lane / 9 - 1;
                if (ax >= 0 && ay >= 0 && az >= 0 && ax < t.ntx && ay < t.nty && az < t.ntz) {
                    tile_enqueue(t, (az * t.nty + ay) * t.ntx + ax, qi_next);
                }
            }
#ifdef SMPLGPU_BFS_STATS
            if (lane == 0) {
                const long long c3 = clock64();
                atomicAdd(reinterpret_cast<unsigned long long*>(t.qn + 8), (unsigned long long)(c1 - c0));
                atomicAdd(reinterpret_cast<unsigned long long*>(t.qn + 10), (unsigned long long)(c2 - c1));
                atomicAdd(reinterpret_cast<unsigned long long*>(t.qn + 12), (unsigned long long)(c3 - c2));
            }
#endif
        }
#ifdef SMPLGPU_BFS_STATS
        const long long b0 = clock64();
#endif
        grid_barrier(bar, (unsigned int)(n + 1) * gridDim.x);
#ifdef SMPLGPU_BFS_STATS
        if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned long long*>(t.qn + 14), (unsigned long long)(clock64() - b0));
#endif
    }
    for (int o = 16; o > 0; o >>= 1) {
        max_level = max(max_level, __shfl_xor_sync(0xffffffffu, max_level, o));
    }
    if (lane == 0 && max_level > 0) {
        atomicMax(&g.ctrl[0], max_level + 1);
    }
}

} // namespace smplgpu
