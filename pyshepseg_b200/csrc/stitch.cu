// stitch.cu -- K9: the device side of tile stitching.
// Replaces the per-pixel work of tiling.recodeTile / recodeSharedSegments / crossesMidline /
// relabelSegments / HistogramAccumulator (tiling.py:1066-1306, 1915-1963).  The reference
// builds a numba Dict of every pixel of every segment twice per tile and loops over the
// segments in Python; here each question the reference asks of those lists becomes a
// per-segment reduction over the label raster:
//   * "top-left corner of the segment's bounding box" (tiling.py:1255-1257)  -> atomicMin of
//     row and column per segment, warp-aggregated;
//   * "does it cross the overlap midline" (tiling.py:1303-1306)              -> min / max of the
//     stitch-axis coordinate over the strip pixels of the segment;
//   * "rank among the segments this tile numbers" (tiling.py:1250-1267)      -> exclusive scan of
//     the numbered flags in ascending id;
//   * "mode of the neighbour's labels under the segment" (tiling.py:1194)    -> (segment,
//     neighbour label) pairs of the crossing segments, radix sorted and run-length encoded;
//     the mode itself is taken on the host after mapping neighbour labels to final ids.
#include "common.cuh"

#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

// What the stitch asks about a segment are all questions of existence -- "is a pixel of it left
// of the trimmed window", "does it have pixels on both sides of the overlap midline" -- so the
// per-segment tables are bytes that any pixel of the segment may set to 1: plain stores, no
// atomics, no read-modify-write.
//   bbox corner inside the trimmed window (tiling.py:1255-1265):
//     minCol >= left  <=> no pixel with c < left        minRow >= top    <=> no pixel with r < top
//     minCol <  right <=> some pixel with c < right      minRow <  bottom <=> some pixel with r < bottom
//   crossesMidline (tiling.py:1303-1306): min < mid <= max over the strip
//     <=> some strip pixel before the midline and some strip pixel at or after it.
struct StitchTables {
    unsigned char *interior;       // has a pixel inside the trimmed window
    unsigned char *margin;         // has a pixel outside it
    unsigned char *leftOf, *above; // has a pixel with c < left / r < top
    unsigned char *ltRight, *ltBottom;   // has a pixel with c < right / r < bottom
    unsigned char *topA, *topB;    // top strip: pixel with r < mid / mid <= r < strip rows
    unsigned char *leftA, *leftB;  // left strip: pixel with c < mid / mid <= c < strip columns
};
#define STITCH_TABLES 10

__global__ void __launch_bounds__(256)
k_tile_max(const unsigned *__restrict__ tile, int64_t N, unsigned long long *counters)
{
    unsigned v = 0;
    const int64_t base = (int64_t)blockIdx.x * blockDim.x * 8 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t p = base + (int64_t)i * blockDim.x;
        if (p < N) v = max(v, tile[p]);
    }
    v = __reduce_max_sync(0xffffffffu, v);
    __shared__ unsigned wmax[8];
    if (lane_id() == 0) wmax[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; w++) v = max(v, wmax[w]);
        if (v) atomicMax(&counters[C_MAXLABEL], (unsigned long long)v);
    }
}

// what one run of equal labels inside one raster row (row r, columns c..cEnd) says about its segment
__device__ __forceinline__ void extents_mark(const StitchTables &tb, unsigned s, unsigned r, unsigned c, unsigned cEnd,
                                             unsigned topRows, unsigned leftCols, unsigned top, unsigned bottom,
                                             unsigned left, unsigned right)
{
    const bool rowInTrim = r >= top && r < bottom;
    const bool touchesTrim = rowInTrim && cEnd >= left && c < right;
    const bool allInTrim = rowInTrim && c >= left && cEnd < right;
    // (plain stores: looking first, to skip a store when the byte is set already, was measured --
    // the byte loads cost twice what the stores do)
    if (touchesTrim) tb.interior[s] = 1;
    if (!allInTrim) {
        // the frame around the trimmed window (a few percent of the pixels)
        tb.margin[s] = 1;
        if (c < left) tb.leftOf[s] = 1;
        if (r < top) tb.above[s] = 1;
        if (c < right) tb.ltRight[s] = 1;
        if (r < bottom) tb.ltBottom[s] = 1;
    }
    if (r < topRows) {
        if (r < topRows / 2) tb.topA[s] = 1; else tb.topB[s] = 1;
    }
    if (c < leftCols) {
        const unsigned mid = leftCols / 2;
        if (c < mid) tb.leftA[s] = 1;
        if (cEnd >= mid) tb.leftB[s] = 1;      // (c < leftCols and cEnd >= mid: a pixel in [mid, leftCols))
    }
}

__global__ void __launch_bounds__(256)
k_tile_extents(const unsigned *__restrict__ tile, int64_t ysize, int64_t xsize, unsigned topRows,
               unsigned leftCols, unsigned top, unsigned bottom, unsigned left, unsigned right,
               StitchTables tb)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < ysize * xsize;
    const unsigned s = valid ? tile[p] : 0u;
    const bool use = valid && s != 0;
    const unsigned r = (unsigned)(p / xsize), c = (unsigned)(p % xsize);
    // a run of equal labels inside one raster row: its head lane knows the row and the first and
    // last column, and speaks for the whole run
    const WarpRuns run = warp_runs(s, use, c == 0);
    if (!run.head) return;
    extents_mark(tb, s, r, c, c + run.len - 1u, topRows, leftCols, top, bottom, left, right);
}

// seg = lut[seg] (the final order-preserving relabel of a tile, shepseg.py:739-777) and, in the
// same pass, the existence tables of the new ids.  A thread takes four consecutive pixels of a
// row (xsize is a multiple of 4 on this path) and speaks for its own runs.
__global__ void __launch_bounds__(256)
k_apply_lut_extents(unsigned *seg, int64_t ysize, int64_t xsize, const unsigned *__restrict__ lut, unsigned topRows,
                    unsigned leftCols, unsigned top, unsigned bottom, unsigned left, unsigned right, StitchTables tb)
{
    const int64_t nGroups = ysize * xsize / 4;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nGroups; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p0 = g * 4;
        uint4 v = *reinterpret_cast<const uint4 *>(seg + p0);
        v.x = __ldg(lut + v.x); v.y = __ldg(lut + v.y); v.z = __ldg(lut + v.z); v.w = __ldg(lut + v.w);
        *reinterpret_cast<uint4 *>(seg + p0) = v;
        const unsigned r = (unsigned)(p0 / xsize), c0 = (unsigned)(p0 - (int64_t)r * xsize);
        const unsigned s[4] = {v.x, v.y, v.z, v.w};
        unsigned start = 0;
#pragma unroll
        for (unsigned i = 0; i < 4; i++) {
            const bool last = i == 3 || s[i + 1 < 4 ? i + 1 : 3] != s[i];
            if (last) {
                if (s[i] != 0) extents_mark(tb, s[i], r, c0 + start, c0 + i, topRows, leftCols, top, bottom, left, right);
                start = i + 1;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_tile_flags(StitchTables tb, int64_t len, int hasTop, int hasLeft, unsigned char *flags,
             unsigned *numbered)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= len) return;
    unsigned f = 0, num = 0;
    const bool interior = tb.interior[s] != 0;
    if (s != 0 && (interior || tb.margin[s])) {
        f |= SSG_SEG_PRESENT;
        if (hasTop && tb.topA[s] && tb.topB[s]) f |= SSG_SEG_KEYTOP;
        if (hasLeft && tb.leftA[s] && tb.leftB[s]) f |= SSG_SEG_KEYLEFT;
        if (interior) f |= SSG_SEG_INTRIM;
        // the corner of the bounding box lies inside the trimmed window (tiling.py:1264-1265);
        // a pixel inside the window is itself left of `right` and above `bottom`
        const bool cornerIn = !tb.leftOf[s] && !tb.above[s] && (interior || tb.ltRight[s]) &&
                              (interior || tb.ltBottom[s]);
        if (!(f & (SSG_SEG_KEYTOP | SSG_SEG_KEYLEFT)) && cornerIn) {
            f |= SSG_SEG_NUMBERED;
            num = 1;
        }
    }
    flags[s] = (unsigned char)f;
    numbered[s] = num;
}

__global__ void __launch_bounds__(256)
k_tile_ranks(const unsigned *__restrict__ numbered, const unsigned *__restrict__ excl,
             const unsigned char *__restrict__ flags, int64_t len, unsigned *rank, unsigned long long *counters)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned inTrim = 0;
    if (s < len) {
        const unsigned r = numbered[s] ? excl[s] + 1u : 0u;
        rank[s] = r;
        if (s == len - 1) counters[C_SCRATCH1] = (unsigned long long)excl[s] + numbered[s];
        if (flags[s] & SSG_SEG_INTRIM) inTrim = r;
    }
    // the highest rank among the numbered segments with a pixel in the trimmed window: by how
    // much the tile moves the running maximum on (tiling.py:1042-1043)
    inTrim = __reduce_max_sync(0xffffffffu, inTrim);
    if ((threadIdx.x & 31) == 0 && inTrim) atomicMax(counters + C_MAXRANKTRIM, (unsigned long long)inTrim);
}

// (strip, segment, neighbour label) keys of the strip pixels of the crossing segments
__global__ void __launch_bounds__(256)
k_collect_pairs(const unsigned *__restrict__ tile, int64_t xsize, int64_t stripRows, int64_t stripCols,
                const unsigned *__restrict__ B, int64_t bStride, const unsigned char *__restrict__ flags,
                unsigned keyFlag, unsigned long long stripBit, unsigned long long *keys,
                unsigned long long *counters)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    unsigned long long key = 0;
    if (t < stripRows * stripCols) {
        const int64_t r = t / stripCols, c = t % stripCols;
        const unsigned s = tile[r * xsize + c];
        if (s != 0 && (flags[s] & keyFlag)) {
            hit = true;
            key = stripBit | ((unsigned long long)s << 32) | (unsigned long long)B[r * bStride + c];
        }
    }
    unsigned long long slot = warp_claim(&counters[C_SCRATCH2], hit);
    if (hit) keys[slot] = key;
}

static void tablesAt(unsigned char *base, size_t n, StitchTables &tb)
{
    tb.interior = base; tb.margin = base + n; tb.leftOf = base + 2 * n; tb.above = base + 3 * n;
    tb.ltRight = base + 4 * n; tb.ltBottom = base + 5 * n; tb.topA = base + 6 * n; tb.topB = base + 7 * n;
    tb.leftA = base + 8 * n; tb.leftB = base + 9 * n;
}

// given: the tables filled by the segmentation (ssg_tile_params.extentsDev), or nullptr
static int reserveTables(ssg_ctx *ctx, int64_t len, StitchTables &tb, unsigned **numbered,
                         unsigned **excl, unsigned **rank, unsigned char **flags,
                         const unsigned char *given, size_t givenStride)
{
    const size_t n = ((size_t)len + 15) & ~(size_t)15;
    SSG_TRY(ssg_reserve(ctx, ctx->stitch1, (size_t)len * 3 * sizeof(unsigned) + (size_t)len));
    if (given) tablesAt(const_cast<unsigned char *>(given), givenStride, tb);
    else {
        SSG_TRY(ssg_reserve(ctx, ctx->stitch0, n * STITCH_TABLES));
        unsigned char *base = bufp<unsigned char>(ctx->stitch0);
        tablesAt(base, n, tb);
        SSG_CUDA(ctx, cudaMemsetAsync(base, 0, n * STITCH_TABLES, ctx->stream));
    }
    unsigned *b1 = bufp<unsigned>(ctx->stitch1);
    *numbered = b1; *excl = b1 + len; *rank = b1 + 2 * len;
    *flags = reinterpret_cast<unsigned char *>(b1 + 3 * len);
    return SSG_OK;
}

// the final relabel of a tile fused with the existence tables (see ssg_tile_params): returns
// false in *done (and does nothing) when the tables do not fit or the vector path does not apply
int ssgk_apply_lut_extents(ssg_ctx *ctx, unsigned *seg, int64_t ysize, int64_t xsize, const unsigned *lut,
                           uint32_t numIds, const ssg_tile_params *prm, uint32_t *stride, bool *done)
{
    *done = false;
    *stride = 0;
    const int64_t len = (int64_t)numIds + 1;
    const size_t n = ((size_t)len + 15) & ~(size_t)15;
    if (!prm->extentsDev || (int64_t)n > prm->extentsCap || xsize % 4 != 0 || ((uintptr_t)seg % 16) != 0 ||
        ysize * xsize == 0)
        return SSG_OK;
    StitchTables tb;
    tablesAt(prm->extentsDev, n, tb);
    SSG_CUDA(ctx, cudaMemsetAsync(prm->extentsDev, 0, n * STITCH_TABLES, ctx->stream));
    int64_t blocks = (ysize * xsize / 4 + 255) / 256;
    if (blocks > (int64_t)ctx->numSMs * 16) blocks = (int64_t)ctx->numSMs * 16;
    SSG_PROF_BEGIN(ctx, "k_apply_lut_extents");
    k_apply_lut_extents<<<(unsigned)blocks, 256, 0, ctx->stream>>>(seg, ysize, xsize, lut, (unsigned)prm->stripRows,
        (unsigned)prm->stripCols, (unsigned)prm->trimTop, (unsigned)prm->trimBottom, (unsigned)prm->trimLeft,
        (unsigned)prm->trimRight, tb);
    SSG_LAUNCHED(ctx);
    *stride = (uint32_t)n;
    *done = true;
    return SSG_OK;
}

extern "C" int ssg_tile_tables_device(ssg_ctx *ctx, const uint32_t *tileDev, int64_t ysize, int64_t xsize,
                                      int64_t overlap, const uint32_t *topBDev, int64_t topBStride,
                                      const uint32_t *leftBDev, int64_t leftBStride, int64_t top,
                                      int64_t bottom, int64_t left, int64_t right, uint32_t maxIdHint,
                                      const uint8_t *extentsDev, int64_t extentsStride, ssg_tile_tables *out)
{
    if (!ctx) return SSG_ERR_ARG;
    ctx->err.clear();
    SSG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!tileDev || !out || ysize <= 0 || xsize <= 0 || overlap < 0) SSG_FAIL(ctx, SSG_ERR_ARG, "bad argument");
    SSG_TRY(ssg_scratch_reset(ctx));
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);
    const int64_t N = ysize * xsize;
    memset(out, 0, sizeof(*out));

    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_MAXLABEL, 0, sizeof(unsigned long long), ctx->stream));
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_SCRATCH1, 0, 2 * sizeof(unsigned long long), ctx->stream));
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_MAXRANKTRIM, 0, sizeof(unsigned long long), ctx->stream));
    unsigned maxId = maxIdHint;
    if (maxId == 0) {       // the caller does not know the largest label: look for it
        SSG_PROF_BEGIN(ctx, "k_tile_max");
        k_tile_max<<<gridFor(N, 256 * 8), 256, 0, ctx->stream>>>(tileDev, N, counters);
        SSG_LAUNCHED(ctx);
        SSG_TRY(ssg_fetch_counters(ctx));
        maxId = (unsigned)ctx->hostCounters[C_MAXLABEL];
    }
    const int64_t len = (int64_t)maxId + 1;

    StitchTables tb;
    unsigned *numbered, *excl, *rank;
    unsigned char *flags;
    if (extentsDev && extentsStride < len) SSG_FAIL(ctx, SSG_ERR_ARG, "extent tables of stride %lld for %lld ids", (long long)extentsStride, (long long)len);
    SSG_TRY(reserveTables(ctx, len, tb, &numbered, &excl, &rank, &flags, extentsDev, (size_t)extentsStride));
    const int64_t topRows = topBDev ? (overlap < ysize ? overlap : ysize) : 0;
    const int64_t leftCols = leftBDev ? (overlap < xsize ? overlap : xsize) : 0;
    if (!extentsDev) {
        SSG_PROF_BEGIN(ctx, "k_tile_extents");
        k_tile_extents<<<gridFor(N, 256), 256, 0, ctx->stream>>>(tileDev, ysize, xsize, (unsigned)topRows, (unsigned)leftCols,
                                                                (unsigned)top, (unsigned)bottom, (unsigned)left,
                                                                (unsigned)right, tb);
        SSG_LAUNCHED(ctx);
    }
    // mid = int(n / 2) of the strip's stitch axis (tiling.py:1297,1300)
    SSG_PROF_BEGIN(ctx, "k_tile_flags");
    k_tile_flags<<<gridFor(len, 256), 256, 0, ctx->stream>>>(tb, len, topBDev != nullptr, leftBDev != nullptr, flags,
                                                            numbered);
    SSG_LAUNCHED(ctx);
    size_t tmpBytes = 0;
    SSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, numbered, excl, (int)len, ctx->stream));
    SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
    SSG_PROF_BEGIN(ctx, "cub_DeviceScan_ExclusiveSum");
    SSG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->cubTemp.p, tmpBytes, numbered, excl, (int)len, ctx->stream));
    SSG_LAUNCHED(ctx);
    SSG_PROF_BEGIN(ctx, "k_tile_ranks");
    k_tile_ranks<<<gridFor(len, 256), 256, 0, ctx->stream>>>(numbered, excl, flags, len, rank, counters);
    SSG_LAUNCHED(ctx);

    // neighbour-label histograms of the crossing segments
    const int64_t nTop = topRows * xsize, nLeft = ysize * leftCols;
    uint32_t numPairs = 0;
    if (nTop + nLeft > 0) {
        SSG_TRY(ssg_reserve(ctx, ctx->stitch2, (size_t)(nTop + nLeft) * sizeof(unsigned long long)));
        unsigned long long *keys = bufp<unsigned long long>(ctx->stitch2);
        if (nTop > 0) {
            SSG_PROF_BEGIN(ctx, "k_collect_pairs");
            k_collect_pairs<<<gridFor(nTop, 256), 256, 0, ctx->stream>>>(tileDev, xsize, topRows, xsize, topBDev, topBStride,
                                                                      flags, SSG_SEG_KEYTOP, 0ull, keys, counters);
            SSG_LAUNCHED(ctx);
        }
        if (nLeft > 0) {
            SSG_PROF_BEGIN(ctx, "k_collect_pairs");
            k_collect_pairs<<<gridFor(nLeft, 256), 256, 0, ctx->stream>>>(tileDev, xsize, ysize, leftCols, leftBDev, leftBStride,
                                                                       flags, SSG_SEG_KEYLEFT, SSG_PAIR_LEFT, keys, counters);
            SSG_LAUNCHED(ctx);
        }
        SSG_TRY(ssg_fetch_counters(ctx));
        const int64_t M = (int64_t)ctx->hostCounters[C_SCRATCH2];
        if (M > 0) {
            SSG_TRY(ssg_reserve(ctx, ctx->stitch3, (size_t)M * sizeof(unsigned long long)));
            SSG_TRY(ssg_reserve(ctx, ctx->stitch4, (size_t)M * sizeof(unsigned long long)));
            SSG_TRY(ssg_reserve(ctx, ctx->stitch5, (size_t)M * sizeof(unsigned)));
            unsigned long long *sorted = bufp<unsigned long long>(ctx->stitch3);
            unsigned long long *uniq = bufp<unsigned long long>(ctx->stitch4);
            unsigned *cnts = bufp<unsigned>(ctx->stitch5);
            SSG_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, keys, sorted, (int)M, 0, 64, ctx->stream));
            SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
            SSG_PROF_BEGIN(ctx, "cub_DeviceRadixSort_SortKeys");
            SSG_CUDA(ctx, cub::DeviceRadixSort::SortKeys(ctx->cubTemp.p, tmpBytes, keys, sorted, (int)M, 0, 64, ctx->stream));
            SSG_LAUNCHED(ctx);
            unsigned long long *dRuns = counters + C_SCRATCH3;
            SSG_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(nullptr, tmpBytes, sorted, uniq, cnts, dRuns, (int)M, ctx->stream));
            SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
            SSG_PROF_BEGIN(ctx, "cub_DeviceRunLengthEncode_Encode");
            SSG_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(ctx->cubTemp.p, tmpBytes, sorted, uniq, cnts, dRuns, (int)M, ctx->stream));
            SSG_LAUNCHED(ctx);
            SSG_TRY(ssg_fetch_counters(ctx));
            numPairs = (uint32_t)(ctx->hostCounters[C_SCRATCH3] & 0xffffffffu);
        }
    } else {
        SSG_TRY(ssg_fetch_counters(ctx));
    }
    out->maxId = maxId;
    out->countNew = (uint32_t)ctx->hostCounters[C_SCRATCH1];
    out->numPairs = numPairs;
    out->maxRankInTrim = (uint32_t)ctx->hostCounters[C_MAXRANKTRIM];
    ctx->stitchLen = len;
    ctx->stitchPairs = numPairs;
    return SSG_OK;
}

extern "C" int ssg_tile_tables_fetch(ssg_ctx *ctx, uint32_t *rank, uint8_t *flags, uint64_t *pairKeys,
                                     uint32_t *pairCounts)
{
    if (!ctx) return SSG_ERR_ARG;
    ctx->err.clear();
    SSG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->stitchLen;
    if (n == 0) SSG_FAIL(ctx, SSG_ERR_STATE, "ssg_tile_tables_device has not been called");
    unsigned *b1 = bufp<unsigned>(ctx->stitch1);
    if (rank) SSG_CUDA(ctx, cudaMemcpyAsync(rank, b1 + 2 * n, n * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    if (flags) SSG_CUDA(ctx, cudaMemcpyAsync(flags, b1 + 3 * n, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->stitchPairs > 0) {
        if (pairKeys) SSG_CUDA(ctx, cudaMemcpyAsync(pairKeys, ctx->stitch4.p, (size_t)ctx->stitchPairs * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (pairCounts) SSG_CUDA(ctx, cudaMemcpyAsync(pairCounts, ctx->stitch5.p, (size_t)ctx->stitchPairs * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    SSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SSG_OK;
}

// out = lut[tile] over the trimmed window (+ histogram of what was written)
__global__ void __launch_bounds__(256)
k_apply_lut_window(const unsigned *__restrict__ tile, int64_t xsize, const unsigned *__restrict__ lut,
                   int64_t top, int64_t left, int64_t wRows, int64_t wCols, unsigned *out,
                   int64_t outStride, unsigned long long *hist, int64_t histLen)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < wRows * wCols;
    unsigned v = 0;
    if (valid) {
        const int64_t r = t / wCols, c = t % wCols;
        v = __ldg(lut + tile[(r + top) * xsize + (c + left)]);
        out[r * outStride + c] = v;
    }
    if (!hist) return;
    const bool use = valid && (int64_t)v < histLen;
    // one atomic per run of equal ids in the warp (a run does not continue into the next row)
    const WarpRuns run = warp_runs(v, use, valid && (t % wCols) == 0);
    if (run.head) atomicAdd(&hist[v], (unsigned long long)run.len);
}

// the same with four consecutive pixels of a row per thread (window width, strides and offsets
// multiples of four; 16-byte loads and stores; one histogram atomic per run inside the thread)
__global__ void __launch_bounds__(256)
k_apply_lut_window4(const unsigned *__restrict__ tile, int64_t xsize, const unsigned *__restrict__ lut,
                    int64_t top, int64_t left, int64_t wRows, int64_t wCols, unsigned *out,
                    int64_t outStride, unsigned long long *hist, int64_t histLen)
{
    const int64_t quadsPerRow = wCols / 4;
    const int64_t nQuads = wRows * quadsPerRow;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nQuads; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = g / quadsPerRow, c = (g - r * quadsPerRow) * 4;
        uint4 v = *reinterpret_cast<const uint4 *>(tile + (r + top) * xsize + (c + left));
        v.x = __ldg(lut + v.x); v.y = __ldg(lut + v.y); v.z = __ldg(lut + v.z); v.w = __ldg(lut + v.w);
        *reinterpret_cast<uint4 *>(out + r * outStride + c) = v;
        if (hist) {
            const unsigned s[4] = {v.x, v.y, v.z, v.w};
            unsigned n = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                n++;
                if (i == 3 || s[i + 1 < 4 ? i + 1 : 3] != s[i]) {
                    if ((int64_t)s[i] < histLen) atomicAdd(&hist[s[i]], (unsigned long long)n);
                    n = 0;
                }
            }
        }
    }
}

// lut = offset + rank where the tile numbered the segment itself (flags == nullptr: every label
// numbers itself, the simple recode), then the ids of the crossing segments over it
__global__ void __launch_bounds__(256)
k_rel_to_lut(unsigned *lut, const unsigned char *__restrict__ flags, int64_t n, unsigned offset)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned v;
    if (flags) v = (flags[i] & SSG_SEG_NUMBERED) ? lut[i] + offset : 0u;
    else v = i ? (unsigned)i + offset : 0u;
    lut[i] = v;
}

__global__ void __launch_bounds__(256)
k_lut_overrides(unsigned *lut, const unsigned *__restrict__ labels, const unsigned *__restrict__ ids, int64_t m)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) lut[labels[i]] = ids[i];
}

static int apply_lut_window(ssg_ctx *ctx, const uint32_t *tileDev, int64_t ysize, int64_t xsize, int64_t top,
                            int64_t bottom, int64_t left, int64_t right, uint32_t *outDev, int64_t outStride,
                            uint64_t *histDev, int64_t histLen);

extern "C" int ssg_apply_lut_device(ssg_ctx *ctx, const uint32_t *tileDev, int64_t ysize, int64_t xsize,
                                    const uint32_t *lutHost, uint32_t maxId, int64_t top, int64_t bottom,
                                    int64_t left, int64_t right, uint32_t *outDev, int64_t outStride,
                                    uint64_t *histDev, int64_t histLen)
{
    if (!ctx) return SSG_ERR_ARG;
    ctx->err.clear();
    SSG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!tileDev || !lutHost || !outDev) SSG_FAIL(ctx, SSG_ERR_ARG, "null pointer argument");
    if (top < 0 || left < 0 || bottom > ysize || right > xsize || top > bottom || left > right) SSG_FAIL(ctx, SSG_ERR_ARG, "bad window");
    SSG_TRY(ssg_scratch_reset(ctx));
    const size_t n = (size_t)maxId + 1;
    SSG_TRY(ssg_reserve(ctx, ctx->lut, n * sizeof(unsigned)));
    ctx->lutStage.assign(lutHost, lutHost + n);
    SSG_CUDA(ctx, cudaMemcpyAsync(ctx->lut.p, ctx->lutStage.data(), n * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    return apply_lut_window(ctx, tileDev, ysize, xsize, top, bottom, left, right, outDev, outStride, histDev, histLen);
}

extern "C" int ssg_apply_rel_lut_device(ssg_ctx *ctx, const uint32_t *tileDev, int64_t ysize, int64_t xsize,
                                        const uint32_t *rankHost, const uint8_t *flagsHost, uint32_t maxId,
                                        uint32_t offset, int64_t nCross,
                                        const uint32_t *crossLabelsHost, const uint32_t *crossIdsHost,
                                        int64_t top, int64_t bottom, int64_t left, int64_t right,
                                        uint32_t *outDev, int64_t outStride, uint64_t *histDev, int64_t histLen)
{
    if (!ctx) return SSG_ERR_ARG;
    ctx->err.clear();
    SSG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!tileDev || !outDev || nCross < 0 || (nCross && (!crossLabelsHost || !crossIdsHost)) || (rankHost && !flagsHost))
        SSG_FAIL(ctx, SSG_ERR_ARG, "null pointer argument");
    if (top < 0 || left < 0 || bottom > ysize || right > xsize || top > bottom || left > right) SSG_FAIL(ctx, SSG_ERR_ARG, "bad window");
    for (int64_t i = 0; i < nCross; i++)
        if (crossLabelsHost[i] > maxId) SSG_FAIL(ctx, SSG_ERR_ARG, "crossing label %u above maxId %u", crossLabelsHost[i], maxId);
    SSG_TRY(ssg_scratch_reset(ctx));
    const int64_t n = (int64_t)maxId + 1;
    const size_t flagWords = rankHost ? ((size_t)n + 3) / 4 : 0;
    SSG_TRY(ssg_reserve(ctx, ctx->lut, ((size_t)n + 2 * (size_t)nCross + flagWords + 2) * sizeof(unsigned)));
    unsigned *lut = bufp<unsigned>(ctx->lut), *crossDev = lut + n, *flagsDev = crossDev + 2 * nCross;
    // one staging vector (it must outlive the asynchronous copies): rank | labels | ids | flags
    const size_t relLen = rankHost ? (size_t)n : 0;
    ctx->lutStage.resize(relLen + 2 * (size_t)nCross + flagWords);
    unsigned *st = ctx->lutStage.data();
    if (rankHost) {
        memcpy(st, rankHost, relLen * sizeof(unsigned));
        memcpy(st + relLen + 2 * nCross, flagsHost, (size_t)n);
    }
    if (nCross) {
        memcpy(st + relLen, crossLabelsHost, (size_t)nCross * sizeof(unsigned));
        memcpy(st + relLen + nCross, crossIdsHost, (size_t)nCross * sizeof(unsigned));
    }
    if (relLen) SSG_CUDA(ctx, cudaMemcpyAsync(lut, st, relLen * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    if (nCross + flagWords) SSG_CUDA(ctx, cudaMemcpyAsync(crossDev, st + relLen, (2 * (size_t)nCross + flagWords) * sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
    SSG_PROF_BEGIN(ctx, "k_rel_to_lut");
    k_rel_to_lut<<<gridFor(n, 256), 256, 0, ctx->stream>>>(lut, rankHost ? reinterpret_cast<const unsigned char *>(flagsDev) : nullptr, n, offset);
    SSG_LAUNCHED(ctx);
    if (nCross) {
        SSG_PROF_BEGIN(ctx, "k_lut_overrides");
        k_lut_overrides<<<gridFor(nCross, 256), 256, 0, ctx->stream>>>(lut, crossDev, crossDev + nCross, nCross);
        SSG_LAUNCHED(ctx);
    }
    return apply_lut_window(ctx, tileDev, ysize, xsize, top, bottom, left, right, outDev, outStride, histDev, histLen);
}

static int apply_lut_window(ssg_ctx *ctx, const uint32_t *tileDev, int64_t ysize, int64_t xsize, int64_t top,
                            int64_t bottom, int64_t left, int64_t right, uint32_t *outDev, int64_t outStride,
                            uint64_t *histDev, int64_t histLen)
{
    const int64_t wRows = bottom - top, wCols = right - left;
    if (wRows * wCols > 0) {
        const bool quads = wCols % 4 == 0 && xsize % 4 == 0 && left % 4 == 0 && outStride % 4 == 0 &&
                           (uintptr_t)tileDev % 16 == 0 && (uintptr_t)outDev % 16 == 0;
        SSG_PROF_BEGIN(ctx, "k_apply_lut_window");
        if (quads) {
            int64_t blocks = (wRows * wCols / 4 + 255) / 256;
            if (blocks > (int64_t)ctx->numSMs * 16) blocks = (int64_t)ctx->numSMs * 16;
            k_apply_lut_window4<<<(unsigned)blocks, 256, 0, ctx->stream>>>(
                tileDev, xsize, bufp<unsigned>(ctx->lut), top, left, wRows, wCols, outDev, outStride,
                reinterpret_cast<unsigned long long *>(histDev), histLen);
        } else {
            k_apply_lut_window<<<gridFor(wRows * wCols, 256), 256, 0, ctx->stream>>>(
                tileDev, xsize, bufp<unsigned>(ctx->lut), top, left, wRows, wCols, outDev, outStride,
                reinterpret_cast<unsigned long long *>(histDev), histLen);
        }
        SSG_LAUNCHED(ctx);
    }
    return SSG_OK;
}

// ---- overviews of a written window (TilingSegmenter.writeOverviews, tiling.py:1360-1383) ------
// For every overview level L the reference keeps every L-th pixel of the window in both
// directions, starting L/2 in from the window's top-left corner: sub = arr[L//2::L, L//2::L].
// All levels of a window are cut out in one launch into one packed buffer (level after level,
// each row-major), which the caller copies to the host in one piece.
#define SSG_MAX_OVERVIEW_LEVELS 16
struct OverviewPlan {
    int n;
    int level[SSG_MAX_OVERVIEW_LEVELS];
    long long rows[SSG_MAX_OVERVIEW_LEVELS], cols[SSG_MAX_OVERVIEW_LEVELS];
    long long start[SSG_MAX_OVERVIEW_LEVELS + 1];       // element offset of every level in the packed buffer
};

__global__ void __launch_bounds__(256)
k_window_overviews(const unsigned *__restrict__ win, int64_t winStride, OverviewPlan plan, unsigned *out)
{
    const long long total = plan.start[plan.n];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int l = 0;
        while (l + 1 < plan.n && i >= plan.start[l + 1]) l++;
        const long long j = i - plan.start[l];
        const long long r = j / plan.cols[l], c = j - r * plan.cols[l];
        const long long L = plan.level[l], o = L / 2;
        out[i] = win[(o + r * L) * winStride + (o + c * L)];
    }
}

extern "C" int ssg_window_overviews(ssg_ctx *ctx, const uint32_t *winDev, int64_t wRows, int64_t wCols,
                                    int64_t winStride, int nLevels, const int32_t *levels, uint32_t *outHost,
                                    int64_t outCapacity, int64_t *startsOut)
{
    if (!ctx) return SSG_ERR_ARG;
    ctx->err.clear();
    SSG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!winDev || !levels || !outHost || !startsOut || nLevels < 1 || nLevels > SSG_MAX_OVERVIEW_LEVELS || wRows < 0 || wCols < 0)
        SSG_FAIL(ctx, SSG_ERR_ARG, "bad argument (1..%d overview levels)", SSG_MAX_OVERVIEW_LEVELS);
    OverviewPlan plan = {};
    plan.n = nLevels;
    for (int l = 0; l < nLevels; l++) {
        if (levels[l] < 1) SSG_FAIL(ctx, SSG_ERR_ARG, "overview level %d", levels[l]);
        const long long L = levels[l], o = L / 2;
        plan.level[l] = levels[l];
        plan.rows[l] = wRows > o ? (wRows - o + L - 1) / L : 0;      // len(range(o, wRows, L))
        plan.cols[l] = wCols > o ? (wCols - o + L - 1) / L : 0;
        plan.start[l + 1] = plan.start[l] + plan.rows[l] * plan.cols[l];
        startsOut[l] = plan.start[l];
    }
    startsOut[nLevels] = plan.start[nLevels];
    const long long total = plan.start[nLevels];
    if (total > outCapacity) SSG_FAIL(ctx, SSG_ERR_ARG, "overview buffer of %lld elements, %lld needed", (long long)outCapacity, total);
    if (total == 0) return SSG_OK;
    SSG_TRY(ssg_reserve(ctx, ctx->stitch2, (size_t)total * sizeof(unsigned)));
    unsigned *out = bufp<unsigned>(ctx->stitch2);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)ctx->numSMs * 16) blocks = (int64_t)ctx->numSMs * 16;
    SSG_PROF_BEGIN(ctx, "k_window_overviews");
    k_window_overviews<<<(unsigned)blocks, 256, 0, ctx->stream>>>(winDev, winStride, plan, out);
    SSG_LAUNCHED(ctx);
    SSG_CUDA(ctx, cudaMemcpyAsync(outHost, out, (size_t)total * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    SSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SSG_OK;
}

