// eliminate.cu -- K4..K8: single-pixel elimination, per-segment spectra, pixel lists and the
// small-segment merge passes, plus the order-preserving relabel.
// Replaces shepseg.eliminateSinglePixels / mergeSinglePixels / findNearestNeighbourPixel
// (shepseg.py:572-736), relabelSegments (739-777), buildSegmentSpectra (780-813),
// makeSegmentLocations (880-915), eliminateSmallSegments / findMergeSegment / doMerge
// (918-1123).
//
// The reference is a chain of sequential loops; what makes it data parallel is that every
// decision phase reads a frozen state (shepseg.py:649-660 and 983-986) and every apply phase
// only depends on order within one target segment (ascending source id, shepseg.py:989-994).
// The kernels keep exactly those two orders and nothing else:
//   * single pixels: per round a decide kernel over the candidates, then an apply kernel;
//   * spectra: exact integer band sums by warp-aggregated 64-bit atomics; a float32 running
//     sum of non-negative integers is exact in any order while the total stays below 2^24,
//     so only segments above that take the ordered float32 chain (raster order, one thread
//     per (segment, band));
//   * pixel lists: a segment's list is a chain of "chunks", one chunk per original segment,
//     each chunk a raster-ordered slice of one pixel array; merging appends the source's
//     chain to the target's (list order = target's pixels, then sources in ascending id,
//     shepseg.py:1099-1114) in O(1);
//   * passes: candidates of the current size are found by a sub-warp each (float32 means,
//     first strict minimum in list order then neighbour order), merges are grouped per
//     target and applied by one thread in ascending source id with float32 adds.
// The whole targetSize loop runs as one persistent cooperative kernel (grid.sync between
// phases) so that ~100 passes do not cost ~500 launches and host round trips.
#include "common.cuh"

#include <cooperative_groups.h>
#include <stdlib.h>
#include <type_traits>
#include <cub/device/device_scan.cuh>
namespace cg = cooperative_groups;

// ====================================================================================
// single pixels
// ====================================================================================
template <typename T>
__device__ __forceinline__ long long pixel_dist(const T *__restrict__ img, int nB, int64_t N,
                                                int64_t p, int64_t q)
{
    unsigned long long d = 0;   // wraps like the reference's int64
    for (int b = 0; b < nB; b++) {
        long long df = (long long)img[(size_t)b * N + p] - (long long)img[(size_t)b * N + q];
        d += (unsigned long long)(df * df);
    }
    return (long long)d;
}

// findNearestNeighbourPixel (shepseg.py:677-736): rows outer, columns inner, first strict
// minimum, only neighbours whose segment has more than one pixel.
template <typename T>
__device__ bool nearest_neighbour(const T *__restrict__ img, int nB, int64_t nRows, int64_t nCols,
                                  const unsigned *__restrict__ seg,
                                  const unsigned *__restrict__ segSize, int four, int64_t p,
                                  unsigned *newSeg)
{
    const int64_t N = nRows * nCols;
    const int64_t i = p / nCols, j = p % nCols;
    long long minD = -1;
    unsigned best = 0;
    bool found = false;
    for (int64_t ii = (i > 0 ? i - 1 : 0); ii <= (i + 1 < nRows ? i + 1 : nRows - 1); ii++) {
        for (int64_t jj = (j > 0 ? j - 1 : 0); jj <= (j + 1 < nCols ? j + 1 : nCols - 1); jj++) {
            if (four && ii != i && jj != j) continue;
            const int64_t q = ii * nCols + jj;
            const unsigned sn = seg[q];
            if (segSize[sn] > 1) {
                long long d = pixel_dist(img, nB, N, p, q);
                if (minD < 0 || d < minD) { minD = d; best = sn; found = true; }
            }
        }
    }
    *newSeg = best;
    return found;
}

// decide phase of mergeSinglePixels (shepseg.py:649-660).  candIn == nullptr: scan every
// pixel (first round); otherwise scan the pixels left over from the previous round.
template <typename T>
__global__ void __launch_bounds__(256)
k_single_decide(const T *__restrict__ img, int nB, int64_t nRows, int64_t nCols,
                const unsigned *__restrict__ seg, const unsigned *__restrict__ segSize, int four,
                const unsigned *__restrict__ candIn, int64_t nIn, unsigned *movePix,
                unsigned *moveSeg, unsigned *candOut, unsigned long long *counters)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool cand = false;
    int64_t p = 0;
    if (t < nIn) {
        p = candIn ? (int64_t)candIn[t] : t;
        cand = segSize[seg[p]] == 1;
    }
    unsigned newSeg = 0;
    bool found = false;
    if (cand) found = nearest_neighbour(img, nB, nRows, nCols, seg, segSize, four, p, &newSeg);
    unsigned long long s0 = warp_claim(&counters[C_NUM_MOVES], cand && found);
    if (cand && found) { movePix[s0] = (unsigned)p; moveSeg[s0] = newSeg; }
    unsigned long long s1 = warp_claim(&counters[C_NUM_LEFT], cand && !found);
    if (cand && !found) candOut[s1] = (unsigned)p;
}

// apply phase (shepseg.py:664-672)
__global__ void __launch_bounds__(256)
k_single_apply(const unsigned *__restrict__ movePix, const unsigned *__restrict__ moveSeg,
               int64_t n, unsigned *seg, unsigned *segSize)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const unsigned p = movePix[t], ns = moveSeg[t];
    const unsigned old = seg[p];
    seg[p] = ns;
    segSize[old] = 0;
    atomicAdd(&segSize[ns], 1u);
}

__global__ void __launch_bounds__(256)
k_count_size_eq(const unsigned *__restrict__ segSize, int64_t lo, int64_t len, unsigned value,
                unsigned long long *counter)
{
    const int64_t s = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = s < len && segSize[s] == value;
    unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane_id() == 0 && m) atomicAdd(counter, (unsigned long long)__popc(m));
}

template <typename T>
static int eliminate_single_t(ssg_ctx *ctx, const T *img, int nB, int64_t nRows, int64_t nCols,
                              unsigned *seg, unsigned *segSize, int64_t len, int four,
                              int64_t *numMoved, unsigned *numRounds)
{
    const int64_t N = nRows * nCols;
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);
    *numMoved = 0;
    *numRounds = 0;
    if (N == 0) return SSG_OK;
    // how many single-pixel segments are there (the lone null pixel counts, shepseg.py:652)
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_SINGLES, 0, sizeof(unsigned long long), ctx->stream));
    SSG_PROF_BEGIN(ctx, "k_count_size_eq");
    k_count_size_eq<<<gridFor(len, 256), 256, 0, ctx->stream>>>(segSize, 0, len, 1u, counters + C_NUM_SINGLES);
    SSG_LAUNCHED(ctx);
    SSG_TRY(ssg_fetch_counters(ctx));
    const int64_t nSingles = (int64_t)ctx->hostCounters[C_NUM_SINGLES];
    if (nSingles == 0) return SSG_OK;
    const size_t cap = (size_t)nSingles * sizeof(unsigned);
    SSG_TRY(ssg_reserve(ctx, ctx->aux0, cap));
    SSG_TRY(ssg_reserve(ctx, ctx->aux1, cap));
    SSG_TRY(ssg_reserve(ctx, ctx->aux2, 2 * cap));
    unsigned *movePix = bufp<unsigned>(ctx->aux0), *moveSeg = bufp<unsigned>(ctx->aux1);
    unsigned *candA = bufp<unsigned>(ctx->aux2), *candB = candA + nSingles;

    const unsigned *candIn = nullptr;
    int64_t nIn = N;
    unsigned *candOut = candA;
    while (true) {
        SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_MOVES, 0, 2 * sizeof(unsigned long long), ctx->stream));
        SSG_PROF_BEGIN(ctx, "k_single_decide");
        k_single_decide<T><<<gridFor(nIn, 256), 256, 0, ctx->stream>>>(img, nB, nRows, nCols, seg, segSize, four,
                                                                       candIn, nIn, movePix, moveSeg, candOut, counters);
        SSG_LAUNCHED(ctx);
        SSG_TRY(ssg_fetch_counters(ctx));
        const int64_t nMoves = (int64_t)ctx->hostCounters[C_NUM_MOVES];
        const int64_t nLeft = (int64_t)ctx->hostCounters[C_NUM_LEFT];
        (*numRounds)++;
        if (nMoves == 0) break;   // the reference's last, empty round (shepseg.py:610)
        SSG_PROF_BEGIN(ctx, "k_single_apply");
        k_single_apply<<<gridFor(nMoves, 256), 256, 0, ctx->stream>>>(movePix, moveSeg, nMoves, seg, segSize);
        SSG_LAUNCHED(ctx);
        *numMoved += nMoves;
        if (nLeft == 0) { (*numRounds)++; break; }   // nothing left to examine: the empty round is implied
        candIn = candOut;
        nIn = nLeft;
        candOut = (candOut == candA) ? candB : candA;
    }
    return SSG_OK;
}

int ssgk_eliminate_single(ssg_ctx *ctx, const void *imgDev, int dtype, int nBands, int64_t nRows,
                          int64_t nCols, uint32_t *segDev, uint32_t *sizeDev, int64_t len,
                          int four, int64_t *numMoved, uint32_t *numRounds)
{
    switch (dtype) {
    case SSG_U8: return eliminate_single_t<uint8_t>(ctx, (const uint8_t *)imgDev, nBands, nRows, nCols, segDev, sizeDev, len, four, numMoved, numRounds);
    case SSG_U16: return eliminate_single_t<uint16_t>(ctx, (const uint16_t *)imgDev, nBands, nRows, nCols, segDev, sizeDev, len, four, numMoved, numRounds);
    case SSG_I16: return eliminate_single_t<int16_t>(ctx, (const int16_t *)imgDev, nBands, nRows, nCols, segDev, sizeDev, len, four, numMoved, numRounds);
    default: SSG_FAIL(ctx, SSG_ERR_ARG, "unsupported dtype code %d", dtype);
    }
}

// ====================================================================================
// relabelSegments (shepseg.py:739-777): new = old - #{j in [minSegId, old-1] : size[j]==0}
// ====================================================================================
__global__ void __launch_bounds__(256)
k_zero_flags(const unsigned *__restrict__ segSize, int64_t len, unsigned minSegId, unsigned *flag)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < len) flag[s] = (s >= minSegId && segSize[s] == 0) ? 1u : 0u;
}

// lut[s] = s - (zeros before s); also the number of ids >= minSegId that own pixels
__global__ void __launch_bounds__(256)
k_make_lut(const unsigned *__restrict__ zerosBefore, const unsigned *__restrict__ segSize,
           int64_t len, unsigned minSegId, unsigned *lut, unsigned long long *counters)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false;
    if (s < len) {
        lut[s] = (unsigned)s - zerosBefore[s];
        alive = s >= minSegId && segSize[s] != 0;
    }
    unsigned m = __ballot_sync(0xffffffffu, alive);
    if (lane_id() == 0 && m) atomicAdd(&counters[C_NUM_ALIVE], (unsigned long long)__popc(m));
}

__global__ void __launch_bounds__(256)
k_apply_lut(unsigned *seg, int64_t N, const unsigned *__restrict__ lut)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p0 = g * 4;
    if (p0 + 4 <= N && ((uintptr_t)seg % 16 == 0)) {
        uint4 v = *reinterpret_cast<uint4 *>(seg + p0);
        v.x = __ldg(lut + v.x); v.y = __ldg(lut + v.y); v.z = __ldg(lut + v.z); v.w = __ldg(lut + v.w);
        *reinterpret_cast<uint4 *>(seg + p0) = v;
    } else {
        for (int64_t p = p0; p < p0 + 4 && p < N; p++) seg[p] = __ldg(lut + seg[p]);
    }
}

int ssgk_relabel(ssg_ctx *ctx, uint32_t *segDev, int64_t N, const uint32_t *sizeDev, int64_t len,
                 uint32_t minSegId, uint32_t *numAlive)
{
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);
    *numAlive = 0;
    if (len <= 0) return SSG_OK;
    SSG_TRY(ssg_reserve(ctx, ctx->lut, (size_t)len * 2 * sizeof(unsigned)));
    unsigned *flag = bufp<unsigned>(ctx->lut), *lut = flag + len;
    SSG_PROF_BEGIN(ctx, "k_zero_flags");
    k_zero_flags<<<gridFor(len, 256), 256, 0, ctx->stream>>>(sizeDev, len, minSegId, flag);
    SSG_LAUNCHED(ctx);
    size_t tmpBytes = 0;
    SSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, flag, flag, (int)len, ctx->stream));
    SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
    SSG_PROF_BEGIN(ctx, "cub_DeviceScan_ExclusiveSum");
    SSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->cubTemp.p, tmpBytes, flag, flag, (int)len, ctx->stream));
    SSG_LAUNCHED(ctx);
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_ALIVE, 0, sizeof(unsigned long long), ctx->stream));
    SSG_PROF_BEGIN(ctx, "k_make_lut");
    k_make_lut<<<gridFor(len, 256), 256, 0, ctx->stream>>>(flag, sizeDev, len, minSegId, lut, counters);
    SSG_LAUNCHED(ctx);
    if (N > 0) {
        SSG_PROF_BEGIN(ctx, "k_apply_lut");
        k_apply_lut<<<gridFor((N + 3) / 4, 256), 256, 0, ctx->stream>>>(segDev, N, lut);
        SSG_LAUNCHED(ctx);
    }
    SSG_TRY(ssg_fetch_counters(ctx));
    *numAlive = (uint32_t)ctx->hostCounters[C_NUM_ALIVE];
    return SSG_OK;
}

// ====================================================================================
// per-segment spectra (buildSegmentSpectra, shepseg.py:780-813)
// ====================================================================================
template <typename T>
__global__ void __launch_bounds__(256)
k_band_sums(const T *__restrict__ img, int nB, int64_t N, const unsigned *__restrict__ seg,
            unsigned long long *isum)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < N;
    const unsigned s = valid ? seg[p] : 0u;
    const bool use = valid && s != 0;
    const unsigned active = __ballot_sync(0xffffffffu, use);
    if (!use) return;
    const unsigned peers = __match_any_sync(active, s);
    const bool leader = (int)lane_id() == __ffs(peers) - 1;
    for (int b = 0; b < nB; b++) {
        int v = (int)img[(size_t)b * N + p];
        int tot = __reduce_add_sync(peers, v);   // <= 32 * 65535, fits
        if (leader) atomicAdd(&isum[(size_t)s * nB + b], (unsigned long long)(long long)tot);
    }
}

// integer sums -> float32 sums where that is exact; flag the segments where it is not
template <bool SIGNED>
__global__ void __launch_bounds__(256)
k_finalize_sums(const unsigned long long *__restrict__ isum, const unsigned *__restrict__ segSize,
                int nB, int64_t len, unsigned maxAbs, float *fsum, unsigned char *bigFlag,
                unsigned long long *counters)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool big = false;
    if (s < len) {
        for (int b = 0; b < nB; b++) {
            long long v = (long long)isum[(size_t)s * nB + b];
            fsum[(size_t)s * nB + b] = (float)v;
            if (!SIGNED) big |= (v >= (1ll << 24));
        }
        if (SIGNED) big = (unsigned long long)segSize[s] * maxAbs >= (1ull << 24);
        if (s == 0) big = false;
        bigFlag[s] = big ? 1 : 0;
    }
    unsigned m = __ballot_sync(0xffffffffu, big);
    if (lane_id() == 0 && m) atomicAdd(&counters[C_NUM_BIGSUM], (unsigned long long)__popc(m));
}

// ordered float32 chain for one (segment, band): s = RN32(s + x) in raster order
template <typename T>
__global__ void __launch_bounds__(128)
k_ordered_sums(const T *__restrict__ img, int nB, int64_t N, const unsigned *__restrict__ pixSorted,
               const unsigned *__restrict__ keysSorted, const unsigned *__restrict__ runStart,
               unsigned numRuns, int64_t M, float *fsum)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t run = t / nB;
    const int b = (int)(t % nB);
    if (run >= numRuns) return;
    const int64_t lo = runStart[run];
    const int64_t hi = (run + 1 < numRuns) ? (int64_t)runStart[run + 1] : M;
    const unsigned s = keysSorted[lo];
    float acc = 0.0f;
    for (int64_t i = lo; i < hi; i++)
        acc = __double2float_rn((double)acc + (double)img[(size_t)b * N + pixSorted[i]]);
    fsum[(size_t)s * nB + b] = acc;
}

template <typename T>
static int build_spectra_t(ssg_ctx *ctx, const T *img, int nB, int64_t N, const unsigned *seg,
                           const unsigned *segSize, int64_t len)
{
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);
    const size_t n = (size_t)len * nB;
    SSG_TRY(ssg_reserve(ctx, ctx->isum, n * sizeof(unsigned long long)));
    SSG_TRY(ssg_reserve(ctx, ctx->fsum, n * sizeof(float)));
    SSG_TRY(ssg_reserve(ctx, ctx->flags, (size_t)len));
    unsigned long long *isum = bufp<unsigned long long>(ctx->isum);
    float *fsum = bufp<float>(ctx->fsum);
    unsigned char *bigFlag = bufp<unsigned char>(ctx->flags);
    SSG_CUDA(ctx, cudaMemsetAsync(isum, 0, n * sizeof(unsigned long long), ctx->stream));
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_BIGSUM, 0, sizeof(unsigned long long), ctx->stream));
    SSG_PROF_BEGIN(ctx, "k_band_sums");
    k_band_sums<T><<<gridFor(N, 256), 256, 0, ctx->stream>>>(img, nB, N, seg, isum);
    SSG_LAUNCHED(ctx);
    constexpr bool isSigned = std::is_signed<T>::value;
    const unsigned maxAbs = sizeof(T) == 1 ? 255u : 32768u;
    SSG_PROF_BEGIN(ctx, "k_finalize_sums");
    k_finalize_sums<isSigned><<<gridFor(len, 256), 256, 0, ctx->stream>>>(isum, segSize, nB, len, maxAbs, fsum, bigFlag, counters);
    SSG_LAUNCHED(ctx);
    SSG_TRY(ssg_fetch_counters(ctx));
    if (ctx->hostCounters[C_NUM_BIGSUM] > 0) {
        const unsigned *pixSorted = nullptr, *keysSorted = nullptr, *runStart = nullptr;
        int64_t M = 0;
        unsigned numRuns = 0;
        SSG_TRY(ssgk_group_pixels(ctx, seg, N, bigFlag, &pixSorted, &keysSorted, &runStart, &M, &numRuns));
        if (M > 0) {
            SSG_PROF_BEGIN(ctx, "k_ordered_sums");
            k_ordered_sums<T><<<gridFor((int64_t)numRuns * nB, 128), 128, 0, ctx->stream>>>(
                img, nB, N, pixSorted, keysSorted, runStart, numRuns, M, fsum);
            SSG_LAUNCHED(ctx);
        }
    }
    return SSG_OK;
}

// ====================================================================================
// pixel lists of the small segments (makeSegmentLocations, shepseg.py:880-915)
// ====================================================================================
__global__ void __launch_bounds__(256)
k_list_counts(const unsigned *__restrict__ segSize, int64_t len, unsigned minSegSize,
              unsigned *cnt /* len + 1 */)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > len) return;
    unsigned c = 0;
    if (s >= 1 && s < len) {
        unsigned z = segSize[s];
        if (z > 0 && z < minSegSize) c = z;
    }
    cnt[s] = c;
}

__global__ void __launch_bounds__(256)
k_list_fill(const unsigned *__restrict__ seg, int64_t N, const unsigned *__restrict__ off,
            unsigned *fill, unsigned *pix)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const unsigned s = seg[p];
    if (s == 0) return;
    const unsigned o = off[s], n = off[s + 1] - o;
    if (n == 0) return;
    const unsigned slot = atomicAdd(&fill[s], 1u);
    pix[o + slot] = (unsigned)p;
}

// slots were claimed in arbitrary order: put every list back into raster order
__global__ void __launch_bounds__(128)
k_list_sort(const unsigned *__restrict__ off, int64_t len, unsigned *pix, unsigned *nextChunk,
            unsigned *tailChunk, unsigned *mergeTo, unsigned *pendHead)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= len) return;
    nextChunk[s] = SSG_NIL;
    tailChunk[s] = (unsigned)s;
    mergeTo[s] = 0;
    pendHead[s] = 0;
    const unsigned o = off[s], n = off[s + 1] - o;
    for (unsigned i = 1; i < n; i++) {
        unsigned v = pix[o + i];
        unsigned j = i;
        while (j > 0 && pix[o + j - 1] > v) { pix[o + j] = pix[o + j - 1]; j--; }
        pix[o + j] = v;
    }
}

// ====================================================================================
// small-segment passes
// ====================================================================================
struct SmallState {
    unsigned *seg;
    unsigned *segSize;
    float *fsum;
    const unsigned *off;      // len+1: pixel-array slice of every original small segment
    const unsigned *pix;
    unsigned *nextChunk, *tailChunk;
    unsigned *mergeTo;
    unsigned *pendHead, *pendNext;
    unsigned *cand, *targets;
    unsigned long long *counters;
    int nB;
    int64_t nRows, nCols;
    int four;
    unsigned len;             // ids are 0..len-1
    int minSegSize;
    double thr;
};

// float32 mean exactly as the reference gets it: float64 division, rounded to float32
// (shepseg.py:1042,1053).  For n < 2^24 a float32 division gives the same bits.
__device__ __forceinline__ float seg_mean(float sum, unsigned n)
{
    if (n < (1u << 24)) return __fdiv_rn(sum, (float)n);
    return __double2float_rn(__ddiv_rn((double)sum, (double)n));
}

__device__ __forceinline__ unsigned group_width(int t)
{
    unsigned g = 1;
    while (g < (unsigned)t && g < 32u) g <<= 1;
    return g;
}

// phase 0: list the segments whose size is exactly t (order irrelevant)
__device__ void phase_enum(const SmallState &st, unsigned t, int par, int64_t gtid, int64_t gsize)
{
    const int64_t n = ((int64_t)st.len + 31) / 32 * 32;   // warp-uniform trip count
    for (int64_t s = gtid; s < n; s += gsize) {
        bool hit = s >= 1 && s < st.len && st.segSize[s] == t;
        unsigned long long slot = warp_claim(&st.counters[C_NUM_CAND0 + par], hit);
        if (hit) st.cand[slot] = (unsigned)s;
    }
}

// phase 1: findMergeSegment (shepseg.py:1003-1063) for every candidate, state frozen.
// A sub-warp of G lanes walks the candidate's chunk chain; lane l takes every G-th pixel of a
// chunk.  The first strict minimum in (list position, neighbour scan order) wins.
template <int NBMAX>
__device__ void phase_find(const SmallState &st, unsigned t, int par, int64_t gtid, int64_t gsize)
{
    const unsigned G = group_width((int)t);
    const unsigned nCand = (unsigned)*(volatile unsigned long long *)&st.counters[C_NUM_CAND0 + par];
    const unsigned lane = lane_id();
    const unsigned sub = lane % G;
    const unsigned groupsPerWarp = 32u / G;
    const int64_t warpId = gtid >> 5, nWarps = gsize >> 5;
    const int nB = st.nB;

    for (int64_t c0 = warpId * groupsPerWarp; c0 < (int64_t)nCand; c0 += nWarps * groupsPerWarp) {
        const int64_t c = c0 + lane / G;
        const bool active = c < (int64_t)nCand;
        unsigned long long bestKey = ~0ull;
        unsigned bestU = 0;
        unsigned s = 0;
        if (active) {
            s = st.cand[c];
            float ms[NBMAX];
#pragma unroll
            for (int b = 0; b < NBMAX; b++)
                if (b < nB) ms[b] = seg_mean(st.fsum[(size_t)s * nB + b], t);
            unsigned posBase = 0;
            unsigned lastU = 0;
            float lastD = 0.0f;
            for (unsigned ch = s; ch != SSG_NIL; ch = st.nextChunk[ch]) {
                const unsigned o = st.off[ch], n = st.off[ch + 1] - o;
                for (unsigned i = sub; i < n; i += G) {
                    const unsigned p = st.pix[o + i];
                    const unsigned k = posBase + i;
                    const int64_t y = p / st.nCols, x = p % st.nCols;
#pragma unroll
                    for (int dy = -1; dy <= 1; dy++) {
                        const int64_t yy = y + dy;
                        if (yy < 0 || yy >= st.nRows) continue;
#pragma unroll
                        for (int dx = -1; dx <= 1; dx++) {
                            const int64_t xx = x + dx;
                            if (xx < 0 || xx >= st.nCols) continue;
                            if (st.four && dy != 0 && dx != 0) continue;
                            const unsigned u = st.seg[yy * st.nCols + xx];
                            if (u == s || u == 0) continue;
                            const unsigned su = st.segSize[u];
                            if (su <= t) continue;
                            float d;
                            if (u == lastU) d = lastD;
                            else {
                                d = 0.0f;
#pragma unroll
                                for (int b = 0; b < NBMAX; b++) {
                                    if (b < nB) {
                                        float mu = seg_mean(st.fsum[(size_t)u * nB + b], su);
                                        float df = __fsub_rn(ms[b], mu);
                                        d = __fadd_rn(d, __fmul_rn(df, df));
                                    }
                                }
                                lastU = u; lastD = d;
                            }
                            const unsigned long long key =
                                ((unsigned long long)__float_as_uint(d) << 32) |
                                (unsigned long long)(k * 16u + (unsigned)((dy + 1) * 3 + (dx + 1)));
                            if (key < bestKey) { bestKey = key; bestU = u; }
                        }
                    }
                }
                posBase += n;
            }
        }
        // minimum over the sub-warp (keys are distinct across lanes: they embed the position)
        for (unsigned o = G >> 1; o > 0; o >>= 1) {
            unsigned long long ok = __shfl_xor_sync(0xffffffffu, bestKey, o);
            unsigned ou = __shfl_xor_sync(0xffffffffu, bestU, o);
            if (ok < bestKey) { bestKey = ok; bestU = ou; }
        }
        if (active && sub == 0 && bestKey != ~0ull) {
            const float d = __uint_as_float((unsigned)(bestKey >> 32));
            if (!((double)d > st.thr)) {       // shepseg.py:1060
                st.mergeTo[s] = bestU;
                const unsigned old = atomicExch(&st.pendHead[bestU], s);
                st.pendNext[s] = old;
                if (old == 0) {
                    unsigned long long slot = atomicAdd(&st.counters[C_NUM_TARGETS0 + par], 1ull);
                    st.targets[slot] = bestU;
                }
            }
        }
    }
}

// phase 2: seg[pixels of source] = target (first half of doMerge, shepseg.py:1107-1110)
__device__ void phase_relabel(const SmallState &st, unsigned t, int par, int64_t gtid, int64_t gsize)
{
    const unsigned G = group_width((int)t);
    const unsigned nCand = (unsigned)*(volatile unsigned long long *)&st.counters[C_NUM_CAND0 + par];
    const unsigned sub = lane_id() % G;
    const int64_t groupId = gtid / G, nGroups = gsize / G;
    for (int64_t c = groupId; c < (int64_t)nCand; c += nGroups) {
        const unsigned s = st.cand[c];
        const unsigned u = st.mergeTo[s];
        if (u == 0) continue;
        for (unsigned ch = s; ch != SSG_NIL; ch = st.nextChunk[ch]) {
            const unsigned o = st.off[ch], n = st.off[ch + 1] - o;
            for (unsigned i = sub; i < n; i += G) st.seg[st.pix[o + i]] = u;
        }
    }
}

// phase 3: per target, its sources in ascending id (shepseg.py:989-994): float32 sum adds,
// size add, chain append (second half of doMerge, shepseg.py:1099-1123)
__device__ void phase_apply(const SmallState &st, int par, int64_t gtid, int64_t gsize)
{
    const unsigned nT = (unsigned)*(volatile unsigned long long *)&st.counters[C_NUM_TARGETS0 + par];
    const int nB = st.nB;
    unsigned long long merged = 0;
    for (int64_t i = gtid; i < (int64_t)nT; i += gsize) {
        const unsigned u = st.targets[i];
        const bool listed = st.off[u + 1] != st.off[u];
        unsigned last = 0;   // ids are >= 1
        while (true) {
            // smallest pending source with id > last
            unsigned nxt = SSG_NIL;
            for (unsigned s = st.pendHead[u]; s != 0; s = st.pendNext[s])
                if (s > last && s < nxt) nxt = s;
            if (nxt == SSG_NIL) break;
            const unsigned s = nxt;
            for (int b = 0; b < nB; b++) {
                float *tu = &st.fsum[(size_t)u * nB + b];
                float *ts = &st.fsum[(size_t)s * nB + b];
                *tu = __fadd_rn(*tu, *ts);
                *ts = 0.0f;
            }
            st.segSize[u] += st.segSize[s];
            st.segSize[s] = 0;
            if (listed) {
                st.nextChunk[st.tailChunk[u]] = s;
                st.tailChunk[u] = st.tailChunk[s];
            }
            st.mergeTo[s] = 0;
            merged++;
            last = s;
        }
        st.pendHead[u] = 0;
    }
    if (merged) atomicAdd(&st.counters[C_NUM_ELIM], merged);
}

// The targetSize loop of eliminateSmallSegments (shepseg.py:970-997) as one cooperative kernel.
template <int NBMAX>
__global__ void __launch_bounds__(256)
k_small_persistent(SmallState st)
{
    cg::grid_group grid = cg::this_grid();
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    int par = 0;
    unsigned long long passes = 0;
    for (int t = 1; t < st.minSegSize; t++) {
        long long prev = -1;
        int numPasses = 0;
        while (true) {
            phase_enum(st, (unsigned)t, par, gtid, gsize);
            grid.sync();
            const long long count = (long long)*(volatile unsigned long long *)&st.counters[C_NUM_CAND0 + par];
            if (gtid == 0) {   // the other parity's counters are idle now: clear them for the next pass
                st.counters[C_NUM_CAND0 + (par ^ 1)] = 0;
                st.counters[C_NUM_TARGETS0 + (par ^ 1)] = 0;
            }
            if (count == prev || numPasses >= 10 || count == 0) { par ^= 1; grid.sync(); break; }
            prev = count;
            phase_find<NBMAX>(st, (unsigned)t, par, gtid, gsize);
            grid.sync();
            phase_relabel(st, (unsigned)t, par, gtid, gsize);
            grid.sync();
            phase_apply(st, par, gtid, gsize);
            par ^= 1;
            numPasses++;
            passes++;
            grid.sync();
        }
    }
    if (gtid == 0) st.counters[C_NUM_PASSES] = passes;
}

// the same phases as separate launches, driven from the host (debugging / comparison)
template <int NBMAX>
__global__ void __launch_bounds__(256) k_phase(SmallState st, int phase, unsigned t, int par)
{
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsize = (int64_t)gridDim.x * blockDim.x;
    if (phase == 0) phase_enum(st, t, par, gtid, gsize);
    else if (phase == 1) phase_find<NBMAX>(st, t, par, gtid, gsize);
    else if (phase == 2) phase_relabel(st, t, par, gtid, gsize);
    else phase_apply(st, par, gtid, gsize);
}

template <int NBMAX>
static int run_small_passes(ssg_ctx *ctx, SmallState &st, uint32_t *numPasses)
{
    unsigned long long *counters = st.counters;
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_ELIM, 0, (C_NUM_PASSES - C_NUM_ELIM + 1) * sizeof(unsigned long long), ctx->stream));
    const char *mode = getenv("SSG_SMALL_MODE");
    const bool hostLoop = mode && strcmp(mode, "host") == 0;
    if (!hostLoop) {
        int perSM = 0;
        SSG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_small_persistent<NBMAX>, 256, 0));
        if (perSM < 1) SSG_FAIL(ctx, SSG_ERR_CUDA, "persistent merge kernel does not fit on an SM");
        if (perSM > 4) perSM = 4;
        dim3 grid((unsigned)(ctx->numSMs * perSM)), block(256);
        void *args[] = {&st};
        SSG_PROF_BEGIN(ctx, "k_small_persistent");
        SSG_CUDA(ctx, cudaLaunchCooperativeKernel((void *)k_small_persistent<NBMAX>, grid, block, args, 0, ctx->stream));
        SSG_LAUNCHED(ctx);
        SSG_TRY(ssg_fetch_counters(ctx));
        *numPasses = (uint32_t)ctx->hostCounters[C_NUM_PASSES];
        return SSG_OK;
    }
    const unsigned grid = (unsigned)ctx->numSMs * 4;
    int par = 0;
    uint32_t passes = 0;
    for (int t = 1; t < st.minSegSize; t++) {
        long long prev = -1;
        int np = 0;
        while (true) {
            SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_CAND0 + par, 0, sizeof(unsigned long long), ctx->stream));
            SSG_CUDA(ctx, cudaMemsetAsync(counters + C_NUM_TARGETS0 + par, 0, sizeof(unsigned long long), ctx->stream));
            SSG_PROF_BEGIN(ctx, "k_phase");
            k_phase<NBMAX><<<grid, 256, 0, ctx->stream>>>(st, 0, (unsigned)t, par);
            SSG_LAUNCHED(ctx);
            SSG_TRY(ssg_fetch_counters(ctx));
            const long long count = (long long)ctx->hostCounters[C_NUM_CAND0 + par];
            if (count == prev || np >= 10 || count == 0) break;
            prev = count;
            for (int ph = 1; ph <= 3; ph++) {
                SSG_PROF_BEGIN(ctx, "k_phase");
                k_phase<NBMAX><<<grid, 256, 0, ctx->stream>>>(st, ph, (unsigned)t, par);
                SSG_LAUNCHED(ctx);
            }
            np++;
            passes++;
        }
    }
    SSG_TRY(ssg_fetch_counters(ctx));
    *numPasses = passes;
    return SSG_OK;
}

template <typename T>
static int eliminate_small_t(ssg_ctx *ctx, const T *img, int nB, int64_t nRows, int64_t nCols,
                             unsigned *seg, unsigned *segSize, uint32_t maxSegId, int minSegSize,
                             double thr, int four, int64_t *numElim, uint32_t *numPasses)
{
    const int64_t N = nRows * nCols;
    const int64_t len = (int64_t)maxSegId + 1;
    *numElim = 0;
    *numPasses = 0;
    if (N == 0 || minSegSize <= 1) return SSG_OK;   // the targetSize loop never runs (shepseg.py:970)
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);

    SSG_TRY(build_spectra_t<T>(ctx, img, nB, N, seg, segSize, len));

    // chunk offsets: exclusive scan of the listed sizes, len+1 entries
    SSG_TRY(ssg_reserve(ctx, ctx->listOff, (size_t)(len + 1) * sizeof(unsigned)));
    unsigned *off = bufp<unsigned>(ctx->listOff);
    SSG_PROF_BEGIN(ctx, "k_list_counts");
    k_list_counts<<<gridFor(len + 1, 256), 256, 0, ctx->stream>>>(segSize, len, (unsigned)minSegSize, off);
    SSG_LAUNCHED(ctx);
    size_t tmpBytes = 0;
    SSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, off, off, (int)(len + 1), ctx->stream));
    SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
    SSG_PROF_BEGIN(ctx, "cub_DeviceScan_ExclusiveSum");
    SSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->cubTemp.p, tmpBytes, off, off, (int)(len + 1), ctx->stream));
    SSG_LAUNCHED(ctx);

    const size_t tbl = (size_t)len * sizeof(unsigned);
    SSG_TRY(ssg_reserve(ctx, ctx->aux0, (size_t)N * sizeof(unsigned)));   // pixel array (<= N entries)
    SSG_TRY(ssg_reserve(ctx, ctx->nextChunk, tbl));
    SSG_TRY(ssg_reserve(ctx, ctx->tailChunk, tbl));
    SSG_TRY(ssg_reserve(ctx, ctx->mergeTo, tbl));
    SSG_TRY(ssg_reserve(ctx, ctx->pendHead, tbl));
    SSG_TRY(ssg_reserve(ctx, ctx->pendNext, tbl));
    SSG_TRY(ssg_reserve(ctx, ctx->candList, tbl));
    SSG_TRY(ssg_reserve(ctx, ctx->targetList, tbl));
    unsigned *pix = bufp<unsigned>(ctx->aux0);
    unsigned *fill = bufp<unsigned>(ctx->candList);   // free until the passes start
    SSG_CUDA(ctx, cudaMemsetAsync(fill, 0, tbl, ctx->stream));
    SSG_PROF_BEGIN(ctx, "k_list_fill");
    k_list_fill<<<gridFor(N, 256), 256, 0, ctx->stream>>>(seg, N, off, fill, pix);
    SSG_LAUNCHED(ctx);
    SSG_PROF_BEGIN(ctx, "k_list_sort");
    k_list_sort<<<gridFor(len, 128), 128, 0, ctx->stream>>>(off, len, pix, bufp<unsigned>(ctx->nextChunk),
                                                           bufp<unsigned>(ctx->tailChunk), bufp<unsigned>(ctx->mergeTo),
                                                           bufp<unsigned>(ctx->pendHead));
    SSG_LAUNCHED(ctx);

    SmallState st;
    st.seg = seg; st.segSize = segSize; st.fsum = bufp<float>(ctx->fsum);
    st.off = off; st.pix = pix;
    st.nextChunk = bufp<unsigned>(ctx->nextChunk); st.tailChunk = bufp<unsigned>(ctx->tailChunk);
    st.mergeTo = bufp<unsigned>(ctx->mergeTo);
    st.pendHead = bufp<unsigned>(ctx->pendHead); st.pendNext = bufp<unsigned>(ctx->pendNext);
    st.cand = bufp<unsigned>(ctx->candList); st.targets = bufp<unsigned>(ctx->targetList);
    st.counters = counters;
    st.nB = nB; st.nRows = nRows; st.nCols = nCols; st.four = four;
    st.len = (unsigned)len; st.minSegSize = minSegSize; st.thr = thr;

    if (nB <= 4) SSG_TRY(run_small_passes<4>(ctx, st, numPasses));
    else if (nB <= 8) SSG_TRY(run_small_passes<8>(ctx, st, numPasses));
    else SSG_TRY(run_small_passes<SSG_MAX_BANDS>(ctx, st, numPasses));
    *numElim = (int64_t)ctx->hostCounters[C_NUM_ELIM];
    return SSG_OK;
}

int ssgk_eliminate_small(ssg_ctx *ctx, const void *imgDev, int dtype, int nBands, int64_t nRows,
                         int64_t nCols, uint32_t *segDev, uint32_t *sizeDev, uint32_t maxSegId,
                         int minSegSize, double thr, int four, int64_t *numElim, uint32_t *numPasses)
{
    switch (dtype) {
    case SSG_U8: return eliminate_small_t<uint8_t>(ctx, (const uint8_t *)imgDev, nBands, nRows, nCols, segDev, sizeDev, maxSegId, minSegSize, thr, four, numElim, numPasses);
    case SSG_U16: return eliminate_small_t<uint16_t>(ctx, (const uint16_t *)imgDev, nBands, nRows, nCols, segDev, sizeDev, maxSegId, minSegSize, thr, four, numElim, numPasses);
    case SSG_I16: return eliminate_small_t<int16_t>(ctx, (const int16_t *)imgDev, nBands, nRows, nCols, segDev, sizeDev, maxSegId, minSegSize, thr, four, numElim, numPasses);
    default: SSG_FAIL(ctx, SSG_ERR_ARG, "unsupported dtype code %d", dtype);
    }
}
