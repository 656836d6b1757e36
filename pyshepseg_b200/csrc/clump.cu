// clump.cu -- K2: connected-component labelling with the reference's numbering.
// Replaces shepseg.clump (shepseg.py:452-541) and shepseg.makeSegSize (544-569).
//
// Result contract: 4-/8-connected regions of equal value (ignoreVal skipped) are numbered
// clumpId, clumpId+1, ... in the raster order of their first pixel; a region of more than
// MAX_CLUMP_SIZE+1 = 10001 pixels is carved into pieces exactly as the reference's capped
// LIFO flood fill carves it (shepseg.py:481,502,523-537).
//
// How it is done here
//   1. ccl_local : 32x64 pixel blocks; horizontal runs from warp ballots, vertical links by
//                  union-find in shared memory (atomicCAS linking to the smaller index),
//                  roots written as global linear indices;
//   2. ccl_border: unions across block borders in global memory;
//   3. ccl_flatten: every pixel points at its root = the raster-first pixel of its region,
//                  which is the reference's seed pixel;
//   4. numbering : roots counted per 1024-pixel block, block offsets by an exclusive scan,
//                  in-block ranks recomputed -> id = clumpId + rank of the root in raster
//                  order; ids gathered to every pixel; sizes by warp-aggregated atomics;
//   5. regions larger than 10001 pixels (none in ordinary imagery) take the slow path: their
//                  pixels are sorted by (region, raster index) and one warp per region replays
//                  the capped flood fill sequentially; its seeds then join the numbering.
#include "common.cuh"

#include <cub/device/device_scan.cuh>

#define CCL_TW 32          // tile width = one warp
#define CCL_TH 64
#define CCL_THREADS 256
#define CCL_ROWS (CCL_TH / (CCL_THREADS / 32))   // rows walked by a warp
#define CCL_PIX (CCL_TW * CCL_TH)
#define SSG_UNVISITED 0xFFFFFFFEu
#define NUM_BLOCK_PIX 1024   // pixels per numbering block (256 threads x 4 consecutive)

// ---- union-find primitives ---------------------------------------------------------------
__device__ __forceinline__ unsigned s_find(volatile unsigned *par, unsigned x)
{
    unsigned p = par[x];
    while (p != x) { x = p; p = par[x]; }
    return x;
}

__device__ __forceinline__ void s_union(unsigned *par, unsigned a, unsigned b)
{
    while (true) {
        a = s_find(par, a);
        b = s_find(par, b);
        if (a == b) return;
        if (a < b) { unsigned t = a; a = b; b = t; }   // link the larger root under the smaller
        unsigned old = atomicCAS(&par[a], a, b);
        if (old == a) return;
        a = old;
    }
}

__device__ __forceinline__ unsigned g_find(const unsigned *L, unsigned x)
{
    unsigned p = __ldcg(L + x);
    while (p != x) { x = p; p = __ldcg(L + x); }
    return x;
}

__device__ __forceinline__ void g_union(unsigned *L, unsigned a, unsigned b)
{
    while (true) {
        a = g_find(L, a);
        b = g_find(L, b);
        if (a == b) return;
        if (a < b) { unsigned t = a; a = b; b = t; }
        unsigned old = atomicCAS(&L[a], a, b);
        if (old == a) return;
        a = old;
    }
}

// ---- 1. block-local labelling --------------------------------------------------------------
// A block labels a 32 x 64 pixel tile; a warp spans the tile's width and walks 8 rows, so a
// pixel's left neighbour is the lane below and the pixel above is the previous iteration's
// register.  Horizontal runs need no union at all: the ballot of "same as the left pixel" gives
// every lane the start of its run by bit arithmetic, and the run start is the pixel's first
// label.  Only the vertical links go through the shared-memory union-find, and of those only
// the first of every stretch where two runs overlap (the rest are implied).
__device__ __forceinline__ unsigned run_start(unsigned eqLeft, unsigned lane)
{
    // highest lane <= mine that does not continue its left neighbour's run
    return 31u - (unsigned)__clz(~eqLeft & (0xffffffffu >> (31u - lane)));
}

__global__ void __launch_bounds__(CCL_THREADS)
k_ccl_local(const int32_t *__restrict__ img, int64_t nRows, int64_t nCols, int32_t ignoreVal,
            int four, unsigned *__restrict__ label)
{
    __shared__ unsigned par[CCL_PIX];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int64_t x0 = (int64_t)blockIdx.x * CCL_TW, y0 = (int64_t)blockIdx.y * CCL_TH;
    const int64_t gx = x0 + lane;
    const bool colOk = gx < nCols;
    const unsigned ly0 = warp * CCL_ROWS;

    // the row above this warp's rows, when it is inside the tile (another warp labels it)
    int32_t upv = ignoreVal;
    unsigned upEq = 0, upRs = lane;
    if (ly0 > 0) {
        const int64_t gy = y0 + ly0 - 1;
        if (colOk && gy < nRows) upv = __ldg(img + gy * nCols + gx);
        const int32_t left = __shfl_up_sync(0xffffffffu, upv, 1);
        upEq = __ballot_sync(0xffffffffu, lane > 0 && upv == left && upv != ignoreVal);
        upRs = run_start(upEq, lane);
    }

    int32_t v[CCL_ROWS];
    unsigned eq[CCL_ROWS], rs[CCL_ROWS];
#pragma unroll
    for (int r = 0; r < CCL_ROWS; r++) {
        const unsigned ly = ly0 + r;
        const int64_t gy = y0 + ly;
        int32_t val = ignoreVal;
        if (colOk && gy < nRows) val = __ldg(img + gy * nCols + gx);
        const int32_t left = __shfl_up_sync(0xffffffffu, val, 1);
        eq[r] = __ballot_sync(0xffffffffu, lane > 0 && val == left && val != ignoreVal);
        rs[r] = run_start(eq[r], lane);
        v[r] = val;
        if (rs[r] == lane) par[ly * CCL_TW + lane] = ly * CCL_TW + lane;
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < CCL_ROWS; r++) {
        const unsigned ly = ly0 + r;
        if (ly == 0) continue;     // the tile's first row: the border kernel links it upwards
        const int32_t up = r == 0 ? upv : v[r > 0 ? r - 1 : 0];
        const unsigned pe = r == 0 ? upEq : eq[r > 0 ? r - 1 : 0];
        const unsigned prs = r == 0 ? upRs : rs[r > 0 ? r - 1 : 0];
        const bool valid = v[r] != ignoreVal;
        const unsigned eu = __ballot_sync(0xffffffffu, valid && v[r] == up);
        const unsigned mine = ly * CCL_TW + rs[r];
        // lane x is implied when x-1 is linked upwards too and both rows continue a run there
        const unsigned need = eu & ~((eu << 1) & eq[r] & pe);
        if ((need >> lane) & 1u) s_union(par, mine, (ly - 1) * CCL_TW + prs);
        if (!four) {
            // diagonals matter only when the orthogonal neighbours do not already connect
            const int32_t ul = __shfl_up_sync(0xffffffffu, up, 1);
            const int32_t ur = __shfl_down_sync(0xffffffffu, up, 1);
            const unsigned prsL = __shfl_up_sync(0xffffffffu, prs, 1);
            const unsigned prsR = __shfl_down_sync(0xffffffffu, prs, 1);
            const bool left = (eq[r] >> lane) & 1u, upSame = (eu >> lane) & 1u;
            if (lane > 0 && valid && v[r] == ul && !left && !upSame)
                s_union(par, mine, (ly - 1) * CCL_TW + prsL);
            if (lane < 31 && valid && v[r] == ur && !upSame)
                s_union(par, mine, (ly - 1) * CCL_TW + prsR);
        }
    }
    __syncthreads();

#pragma unroll
    for (int r = 0; r < CCL_ROWS; r++) {
        const unsigned ly = ly0 + r;
        const int64_t gy = y0 + ly;
        if (!colOk || gy >= nRows) continue;
        unsigned out = SSG_NIL;
        if (v[r] != ignoreVal) {
            const unsigned root = s_find(par, ly * CCL_TW + rs[r]);
            out = (unsigned)((y0 + root / CCL_TW) * nCols + x0 + (root % CCL_TW));
        }
        label[gy * nCols + gx] = out;
    }
}

// ---- 2. unions across block borders ----------------------------------------------------------
__global__ void __launch_bounds__(256)
k_ccl_border(const int32_t *__restrict__ img, int64_t nRows, int64_t nCols, int32_t ignoreVal,
             int four, unsigned *label, int64_t nHB, int64_t nVB)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nH = nHB * nCols;
    int64_t y, x;
    bool horiz;
    if (t < nH) { y = (t / nCols + 1) * CCL_TH; x = t % nCols; horiz = true; }
    else {
        int64_t u = t - nH;
        if (u >= nVB * nRows) return;
        x = (u / nRows + 1) * CCL_TW; y = u % nRows; horiz = false;
    }
    const int64_t p = y * nCols + x;
    const int32_t v = __ldg(img + p);
    if (v == ignoreVal) return;
    if (horiz) {
        // neighbours in the row above (other block row)
        if (__ldg(img + p - nCols) == v) g_union(label, (unsigned)p, (unsigned)(p - nCols));
        if (!four) {
            if (x > 0 && __ldg(img + p - nCols - 1) == v) g_union(label, (unsigned)p, (unsigned)(p - nCols - 1));
            if (x < nCols - 1 && __ldg(img + p - nCols + 1) == v) g_union(label, (unsigned)p, (unsigned)(p - nCols + 1));
        }
    } else {
        // neighbours in the column to the left (other block column)
        if (__ldg(img + p - 1) == v) g_union(label, (unsigned)p, (unsigned)(p - 1));
        if (!four) {
            if (y > 0 && __ldg(img + p - nCols - 1) == v) g_union(label, (unsigned)p, (unsigned)(p - nCols - 1));
            if (y < nRows - 1 && __ldg(img + p + nCols - 1) == v) g_union(label, (unsigned)p, (unsigned)(p + nCols - 1));
        }
    }
}

// ---- 3. flatten + 4a. count the roots ----------------------------------------------------------
// every pixel points at its root; a pixel is a root (the seed of a clump) iff it points at
// itself, which the flattening does not change, so the per-block root counts of the numbering
// are taken in the same pass (thread = 4 consecutive pixels, block = NUM_BLOCK_PIX)
__global__ void __launch_bounds__(256)
k_ccl_flatten_count(unsigned *label, int64_t N, unsigned *__restrict__ blockCnt,
                    unsigned long long *counters)
{
    const int64_t base = (int64_t)blockIdx.x * NUM_BLOCK_PIX + (int64_t)threadIdx.x * 4;
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t p = base + i;
        if (p >= N) break;
        const unsigned l = label[p];
        if (l == SSG_NIL) continue;
        if (l == (unsigned)p) { cnt++; continue; }
        const unsigned r = g_find(label, l);
        if (r != l) label[p] = r;
    }
    __shared__ int wsum[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane_id() == 0) wsum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) tot += wsum[w];
        blockCnt[blockIdx.x] = (unsigned)tot;
        if (tot) atomicAdd(&counters[C_NUM_ROOTS], (unsigned long long)tot);
    }
}

__global__ void __launch_bounds__(256)
k_number_roots(const unsigned *__restrict__ label, int64_t N, const unsigned *__restrict__ blockOff,
               unsigned clumpId, unsigned *__restrict__ seg)
{
    const int64_t base = (int64_t)blockIdx.x * NUM_BLOCK_PIX + (int64_t)threadIdx.x * 4;
    bool isRoot[4];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int64_t p = base + i;
        isRoot[i] = (p < N && label[p] == (unsigned)p);
        cnt += isRoot[i];
    }
    // exclusive prefix of cnt over the block (raster order = thread order)
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane_id() >= o) incl += v;
    }
    __shared__ int wsum[8];
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) wbase += wsum[w];
    unsigned id = clumpId + blockOff[blockIdx.x] + (unsigned)(wbase + incl - cnt);
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (isRoot[i]) seg[base + i] = id++;
}

// ids of the roots to every pixel + the size table.  A root whose later neighbours (right and
// below; a root is the raster-first pixel of its clump) do not point at it is a clump of one
// pixel: those are listed for the single-pixel stage, which then needs no scan of its own.
// A thread takes four consecutive pixels (16-byte loads and stores); sizes go up by one atomic per
// run of equal ids inside the thread, a single-pixel clump by a plain store.
#define GATHER_PIX 1024   // pixels per block (256 threads x 4)
__global__ void __launch_bounds__(256)
k_gather_ids(const unsigned *__restrict__ label, int64_t nRows, int64_t nCols, int four,
             unsigned *seg, unsigned *segSize, unsigned *singles, unsigned long long *counters)
{
    __shared__ unsigned sList[GATHER_PIX];
    __shared__ unsigned sCount;
    __shared__ unsigned long long sBase;
    const int64_t N = nRows * nCols;
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();
    const int64_t p0 = (int64_t)blockIdx.x * GATHER_PIX + (int64_t)threadIdx.x * 4;
    unsigned l[4], id[4];
    bool single[4] = {false, false, false, false};
    const bool vec = p0 + 4 <= N && (N % 4 == 0);
    if (vec) {
        const uint4 lv = *reinterpret_cast<const uint4 *>(label + p0);
        l[0] = lv.x; l[1] = lv.y; l[2] = lv.z; l[3] = lv.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) l[i] = p0 + i < N ? label[p0 + i] : SSG_NIL;
    }
    unsigned nextLabel = SSG_NIL;      // label of the pixel after my four (same row or not: checked below)
    bool anyRoot = false;
#pragma unroll
    for (int i = 0; i < 4; i++) anyRoot |= (p0 + i < N && l[i] == (unsigned)(p0 + i));
    if (anyRoot && p0 + 4 < N) nextLabel = label[p0 + 4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t p = p0 + i;
        id[i] = 0;
        if (p >= N || l[i] == SSG_NIL) continue;
        if (l[i] == (unsigned)p) {
            id[i] = seg[p];            // written by k_number_roots (previous launch)
            if (singles) {
                const int64_t y = p / nCols, x = p - y * nCols;
                const bool hasR = x + 1 < nCols, hasD = y + 1 < nRows;
                const unsigned right = i < 3 ? l[i + 1 < 4 ? i + 1 : 3] : nextLabel;
                bool grown = (hasR && right == l[i]) || (hasD && label[p + nCols] == l[i]);
                if (!four && hasD)
                    grown = grown || (x > 0 && label[p + nCols - 1] == l[i]) || (hasR && label[p + nCols + 1] == l[i]);
                single[i] = !grown;
            }
        } else {
            id[i] = __ldcg(seg + l[i]);
        }
    }
    if (vec) *reinterpret_cast<uint4 *>(seg + p0) = make_uint4(id[0], id[1], id[2], id[3]);
    else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (p0 + i < N) seg[p0 + i] = id[i];
    }
    // sizes: one update per run of equal ids among my four pixels
    unsigned n = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (p0 + i >= N) break;
        n++;
        const bool last = i == 3 || p0 + i + 1 >= N || id[i + 1 < 4 ? i + 1 : 3] != id[i];
        if (last) {
            if (single[i]) segSize[id[i]] = 1u;      // nobody else touches a one-pixel clump's size
            else atomicAdd(&segSize[id[i]], n);
            n = 0;
        }
    }
    if (!singles) return;
    // block-local list of the single-pixel clumps first: one global atomic per block
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const unsigned m = __ballot_sync(0xffffffffu, single[i]);
        if (m) {
            unsigned base = 0;
            const int leader = __ffs(m) - 1;
            if ((int)lane_id() == leader) base = atomicAdd(&sCount, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (single[i]) sList[base + __popc(m & ((1u << lane_id()) - 1u))] = (unsigned)(p0 + i);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sCount) sBase = atomicAdd(&counters[C_NUM_SINGLEPIX], (unsigned long long)sCount);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < sCount; i += 256) singles[sBase + i] = sList[i];
}

__global__ void __launch_bounds__(256)
k_count_oversized(const unsigned *__restrict__ segSize, int64_t lo, int64_t len,
                  unsigned long long *counters)
{
    const int64_t s = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool over = s < len && segSize[s] > SSG_MAX_CLUMP_SIZE + 1;
    const bool single = s < len && segSize[s] == 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) counters[C_NULL_SINGLE] = segSize[0] == 1 ? 1ull : 0ull;
    block_add(&counters[C_NUM_OVERSIZED], over ? 1u : 0u);
    block_add(&counters[C_NUM_SINGLES], single ? 1u : 0u);
}

// labels (root pointers) -> dense ids in seg + size table; returns the number of roots
static int number_from_labels(ssg_ctx *ctx, unsigned *label, int64_t nRows, int64_t nCols,
                              int four, unsigned clumpId, unsigned *seg, unsigned *numRoots,
                              bool listSingles)
{
    const int64_t N = nRows * nCols;
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);
    const int64_t nBlocks = (N + NUM_BLOCK_PIX - 1) / NUM_BLOCK_PIX;
    SSG_TRY(ssg_reserve(ctx, ctx->blockCnt, (size_t)nBlocks * 2 * sizeof(unsigned)));
    unsigned *blockCnt = bufp<unsigned>(ctx->blockCnt);
    unsigned *blockOff = blockCnt + nBlocks;
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_ROOTS, 0, sizeof(unsigned long long), ctx->stream));
    SSG_PROF_BEGIN(ctx, "k_ccl_flatten_count");
    k_ccl_flatten_count<<<(unsigned)nBlocks, 256, 0, ctx->stream>>>(label, N, blockCnt, counters);
    SSG_LAUNCHED(ctx);
    size_t tmpBytes = 0;
    SSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, blockCnt, blockOff, (int)nBlocks, ctx->stream));
    SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
    SSG_PROF_BEGIN(ctx, "cub_DeviceScan_ExclusiveSum");
    SSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->cubTemp.p, tmpBytes, blockCnt, blockOff, (int)nBlocks, ctx->stream));
    SSG_LAUNCHED(ctx);
    SSG_PROF_BEGIN(ctx, "k_number_roots");
    k_number_roots<<<(unsigned)nBlocks, 256, 0, ctx->stream>>>(label, N, blockOff, clumpId, seg);
    SSG_LAUNCHED(ctx);
    SSG_TRY(ssg_fetch_counters(ctx));
    *numRoots = (unsigned)ctx->hostCounters[C_NUM_ROOTS];
    const size_t len = (size_t)clumpId + *numRoots;
    SSG_TRY(ssg_reserve(ctx, ctx->segSize, len * sizeof(unsigned)));
    SSG_CUDA(ctx, cudaMemsetAsync(ctx->segSize.p, 0, len * sizeof(unsigned), ctx->stream));
    unsigned *singles = nullptr;
    if (listSingles) {
        SSG_TRY(ssg_reserve(ctx, ctx->singles, (size_t)N * sizeof(unsigned)));
        singles = bufp<unsigned>(ctx->singles);
        SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_SINGLEPIX, 0, sizeof(unsigned long long), ctx->stream));
    }
    SSG_PROF_BEGIN(ctx, "k_gather_ids");
    k_gather_ids<<<gridFor(N, GATHER_PIX), 256, 0, ctx->stream>>>(label, nRows, nCols, four, seg, bufp<unsigned>(ctx->segSize),
                                                          singles, counters);
    SSG_LAUNCHED(ctx);
    return SSG_OK;
}

// ---- 5. slow path: regions over the cap -------------------------------------------------------
__global__ void __launch_bounds__(256)
k_flag_oversized(const unsigned *__restrict__ segSize, int64_t len, unsigned char *flag)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < len) flag[s] = (s != 0 && segSize[s] > SSG_MAX_CLUMP_SIZE + 1) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_mark_unvisited(const unsigned *__restrict__ pix, int64_t M, unsigned *label)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) label[pix[i]] = SSG_UNVISITED;
}

// One warp replays the reference's capped flood fill over one region (shepseg.py:490-539).
// Lane l looks at window cell (cx, cy) = (sx-1 + l/3, sy-1 + l%3): columns outer, rows inner,
// which is the order the reference pushes neighbours in (shepseg.py:523-524).
// (one warp per block, the LIFO stack in shared memory: the fill is a chain of dependent steps --
// pop, look at the 3x3 window, push -- and a stack in global memory doubled the length of each)
__global__ void __launch_bounds__(32)
k_capped_fill(const int32_t *__restrict__ img, int64_t nRows, int64_t nCols, int four,
              const unsigned *__restrict__ sortedPix, const unsigned *__restrict__ runStart,
              unsigned numRuns, int64_t M, unsigned *label)
{
    __shared__ unsigned stack[SSG_MAX_CLUMP_SIZE + 32];
    const unsigned warp = blockIdx.x;
    if (warp >= numRuns) return;
    const unsigned lane = lane_id();
    const int64_t lo = runStart[warp];
    const int64_t hi = (warp + 1 < numRuns) ? (int64_t)runStart[warp + 1] : M;
    volatile unsigned *vlabel = label;
    const int dx = (int)(lane / 3) - 1, dy = (int)(lane % 3) - 1;
    const bool cellUsed = lane < 9 && (!four || dx == 0 || dy == 0);

    for (int64_t b0 = lo; b0 < hi; b0 += 32) {
        const int64_t i = b0 + lane;
        const unsigned myPix = i < hi ? sortedPix[i] : 0u;
        while (true) {
            bool unvis = i < hi && vlabel[myPix] == SSG_UNVISITED;
            unsigned m = __ballot_sync(0xffffffffu, unvis);
            if (m == 0) break;
            const int src = __ffs(m) - 1;
            const unsigned seed = __shfl_sync(0xffffffffu, myPix, src);
            const int32_t val = __ldg(img + seed);
            unsigned sp = 1, n = 0;
            if (lane == 0) { vlabel[seed] = seed; stack[0] = seed; }
            __syncwarp();
            while (sp > 0 && n < SSG_MAX_CLUMP_SIZE) {
                sp--;
                const unsigned q = ((volatile unsigned *)stack)[sp];
                const int64_t sy = q / nCols, sx = q % nCols;
                const int64_t cx = sx + dx, cy = sy + dy;
                bool ok = cellUsed && cx >= 0 && cx < nCols && cy >= 0 && cy < nRows;
                unsigned nb = 0;
                if (ok) {
                    nb = (unsigned)(cy * nCols + cx);
                    ok = __ldg(img + nb) == val && vlabel[nb] == SSG_UNVISITED;
                }
                const unsigned pm = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    vlabel[nb] = seed;
                    ((volatile unsigned *)stack)[sp + __popc(pm & ((1u << lane) - 1u))] = nb;
                }
                const unsigned cnt = __popc(pm);
                sp += cnt;
                n += cnt;
                __syncwarp();
            }
            __syncwarp();
        }
    }
}

static int split_oversized(ssg_ctx *ctx, const int32_t *img, int64_t nRows, int64_t nCols, int four,
                           unsigned *label, const unsigned *seg, int64_t sizeLen)
{
    const int64_t N = nRows * nCols;
    SSG_TRY(ssg_reserve(ctx, ctx->flags, (size_t)sizeLen));
    unsigned char *flag = bufp<unsigned char>(ctx->flags);
    SSG_PROF_BEGIN(ctx, "k_flag_oversized");
    k_flag_oversized<<<gridFor(sizeLen, 256), 256, 0, ctx->stream>>>(bufp<unsigned>(ctx->segSize), sizeLen, flag);
    SSG_LAUNCHED(ctx);
    // pixels of the flagged regions, grouped by region, raster order inside each region
    const unsigned *pixSorted = nullptr, *runStart = nullptr;
    int64_t M = 0;
    unsigned numRuns = 0;
    SSG_TRY(ssgk_group_pixels(ctx, seg, N, flag, &pixSorted, nullptr, &runStart, &M, &numRuns));
    if (M == 0) return SSG_OK;
    SSG_PROF_BEGIN(ctx, "k_mark_unvisited");
    k_mark_unvisited<<<gridFor(M, 256), 256, 0, ctx->stream>>>(pixSorted, M, label);
    SSG_LAUNCHED(ctx);
    SSG_PROF_BEGIN(ctx, "k_capped_fill");
    k_capped_fill<<<numRuns, 32, 0, ctx->stream>>>(img, nRows, nCols, four, pixSorted, runStart, numRuns, M, label);
    SSG_LAUNCHED(ctx);
    return SSG_OK;
}

// ---- driver ----------------------------------------------------------------------------------
int ssgk_clump(ssg_ctx *ctx, const int32_t *clusterDev, int64_t nRows, int64_t nCols,
               int32_t ignoreVal, int four, uint32_t clumpId, uint32_t *segDev,
               uint32_t *numClumps, uint32_t *numOversized, int64_t *singlesOut)
{
    const int64_t N = nRows * nCols;
    *numClumps = 0;
    *numOversized = 0;
    if (N == 0) return SSG_OK;
    if (N >= 0xFFFFFFF0ll) SSG_FAIL(ctx, SSG_ERR_ARG, "tile of %lld pixels is too large for 32-bit pixel indices", (long long)N);
    SSG_TRY(ssg_reserve(ctx, ctx->label, (size_t)N * sizeof(unsigned)));
    unsigned *label = bufp<unsigned>(ctx->label);
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);

    dim3 grid((unsigned)((nCols + CCL_TW - 1) / CCL_TW), (unsigned)((nRows + CCL_TH - 1) / CCL_TH));
    SSG_PROF_BEGIN(ctx, "k_ccl_local");
    k_ccl_local<<<grid, CCL_THREADS, 0, ctx->stream>>>(clusterDev, nRows, nCols, ignoreVal, four, label);
    SSG_LAUNCHED(ctx);
    const int64_t nHB = (nRows - 1) / CCL_TH, nVB = (nCols - 1) / CCL_TW;
    const int64_t nBorder = nHB * nCols + nVB * nRows;
    if (nBorder > 0) {
        SSG_PROF_BEGIN(ctx, "k_ccl_border");
        k_ccl_border<<<gridFor(nBorder, 256), 256, 0, ctx->stream>>>(clusterDev, nRows, nCols, ignoreVal, four, label, nHB, nVB);
        SSG_LAUNCHED(ctx);
    }
    unsigned numRoots = 0;
    SSG_TRY(number_from_labels(ctx, label, nRows, nCols, four, clumpId, segDev, &numRoots, singlesOut != nullptr));
    // oversized regions / single-pixel clumps
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_OVERSIZED, 0, 2 * sizeof(unsigned long long), ctx->stream));
    SSG_PROF_BEGIN(ctx, "k_count_oversized");
    k_count_oversized<<<gridFor(numRoots, 256), 256, 0, ctx->stream>>>(bufp<unsigned>(ctx->segSize), (int64_t)clumpId,
                                                                       (int64_t)clumpId + numRoots, counters);
    SSG_LAUNCHED(ctx);
    SSG_TRY(ssg_fetch_counters(ctx));
    const unsigned nOver = (unsigned)ctx->hostCounters[C_NUM_OVERSIZED];
    *numOversized = nOver;
    if (nOver > 0) {
        SSG_TRY(split_oversized(ctx, clusterDev, nRows, nCols, four, label, segDev, (int64_t)clumpId + numRoots));
        SSG_TRY(number_from_labels(ctx, label, nRows, nCols, four, clumpId, segDev, &numRoots, singlesOut != nullptr));
        SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_OVERSIZED, 0, 2 * sizeof(unsigned long long), ctx->stream));
        SSG_PROF_BEGIN(ctx, "k_count_oversized");
        k_count_oversized<<<gridFor(numRoots, 256), 256, 0, ctx->stream>>>(bufp<unsigned>(ctx->segSize), (int64_t)clumpId,
                                                                           (int64_t)clumpId + numRoots, counters);
        SSG_LAUNCHED(ctx);
        SSG_TRY(ssg_fetch_counters(ctx));
    }
    *numClumps = numRoots;
    if (singlesOut) {
        // the list stands for "every pixel whose segment has one pixel" unless the null segment
        // is such a pixel too (shepseg.py:652 treats it like any other segment)
        const int64_t listed = (int64_t)ctx->hostCounters[C_NUM_SINGLEPIX];
        const bool ok = ctx->hostCounters[C_NULL_SINGLE] == 0 && listed == (int64_t)ctx->hostCounters[C_NUM_SINGLES];
        *singlesOut = ok ? listed : -1;
    }
    return SSG_OK;
}

// ---- makeSegSize on an arbitrary label raster (shepseg.py:544-569) -----------------------------
__global__ void __launch_bounds__(256)
k_seg_size(const unsigned *__restrict__ seg, int64_t N, unsigned *segSize, int64_t len)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = p < N;
    unsigned id = valid ? seg[p] : 0;
    if (valid && (int64_t)id >= len) valid = false;
    const WarpRuns run = warp_runs(id, valid);
    if (run.head) atomicAdd(&segSize[id], run.len);
}

int ssgk_seg_size(ssg_ctx *ctx, const uint32_t *segDev, int64_t N, uint32_t *sizeDev, int64_t len)
{
    SSG_CUDA(ctx, cudaMemsetAsync(sizeDev, 0, (size_t)len * sizeof(unsigned), ctx->stream));
    if (N == 0) return SSG_OK;
    SSG_PROF_BEGIN(ctx, "k_seg_size");
    k_seg_size<<<gridFor(N, 256), 256, 0, ctx->stream>>>(segDev, N, sizeDev, len);
    SSG_LAUNCHED(ctx);
    return SSG_OK;
}
