// stats.cu -- per-segment statistics of an image band over a segment raster.
// Replaces the accumulation and the SegmentStats class of pyshepseg.tilingstats
// (accumulateSegDict, tilingstats.py:467-517; SegmentStats, 923-1008; calcStatsForCompletedSegs,
// 557-617): the reference builds, tile by tile, a dictionary of value histograms per segment and
// evaluates min / max / mean / stddev / median / mode / percentile / pixcount from the histogram
// sorted by value.  Here the (segment, value) pairs of all valid pixels are sorted once (radix
// sort of packed keys), run-length encoded into exactly those sorted histograms, and one thread
// per segment evaluates the statistics with the reference's arithmetic:
//   mean    int64 sum of value*count / pixCount in float64, stored as float32 (tilingstats.py:953);
//   stddev  float32 sum, in ascending value, of float32(count * (value - float32 mean)^2), / pixCount, sqrt,
//           stored as float32 (956-957);
//   mode    the smallest value among those with the highest count (numpy.argmax, 960);
//   percentile p: the value at which the running count first reaches pixCount * (p / 100) (969-986);
//   a segment without valid pixels gets missingStatsValue in every column but pixcount, which is 0.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_select.cuh>

#define STAT_MIN 0
#define STAT_MAX 1
#define STAT_MEAN 2
#define STAT_STDDEV 3
#define STAT_MEDIAN 4
#define STAT_MODE 5
#define STAT_PERCENTILE 6
#define STAT_PIXCOUNT 7

struct IsValidKey {
    __device__ bool operator()(unsigned long long k) const { return k != ~0ull; }
};

// key = segment << valueBits | (value - valueMin); ~0 for pixels that do not count
template <typename T>
__global__ void __launch_bounds__(256)
k_stats_keys(const unsigned *__restrict__ seg, const T *__restrict__ img, int64_t N, int hasNull, long long nullVal,
             long long valueMin, int valueBits, unsigned maxSegId, unsigned *nullCount, unsigned long long *outOfRange,
             unsigned long long *keys)
{
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (int64_t)gridDim.x * blockDim.x) {
        const unsigned s = seg[p];
        const long long v = (long long)img[p];
        unsigned long long k = ~0ull;
        if (s > maxSegId) atomicAdd(outOfRange, 1ull);
        else if (s != 0) {
            if (hasNull && v == nullVal) atomicAdd(nullCount + s, 1u);     // (the reference's noDataDict)
            else k = ((unsigned long long)s << valueBits) | (unsigned long long)(v - valueMin);
        }
        keys[p] = k;
    }
}

// first run of every segment (runs are sorted by segment, then value)
__global__ void __launch_bounds__(256)
k_stats_starts(const unsigned long long *__restrict__ runKeys, int64_t nRuns, int valueBits, unsigned *segStart)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nRuns) return;
    const unsigned s = (unsigned)(runKeys[i] >> valueBits);
    if (i == 0 || (unsigned)(runKeys[i - 1] >> valueBits) != s) segStart[s] = (unsigned)i;
}

struct StatSel {
    int statId[32];
    int param[32];
    int column[32];     // index among the int64 / float32 output columns
    int isFloat[32];
    int n;
};

__global__ void __launch_bounds__(128)
k_stats_eval(const unsigned long long *__restrict__ runKeys, const unsigned *__restrict__ runCounts, int64_t nRuns,
             const unsigned *__restrict__ segStart, int64_t len, int valueBits, long long valueMin,
             long long missing, StatSel sel, const unsigned *__restrict__ nullCount, long long *intOut, float *floatOut,
             unsigned *totalOut)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= len) return;
    if (s == 0) {      // the null segment's row is zero (RatPage, tilingstats.py:1990-1994)
        for (int i = 0; i < sel.n; i++) {
            if (sel.isFloat[i]) floatOut[(size_t)sel.column[i] * len] = 0.0f;
            else intOut[(size_t)sel.column[i] * len] = 0;
        }
        totalOut[0] = 0;
        return;
    }
    const unsigned long long mask = (1ull << valueBits) - 1ull;
    const unsigned first = segStart[s];
    unsigned long long pixCount = 0;
    long long vmin = 0, vmax = 0, mode = 0, sumVC = 0;
    unsigned last = first;
    if (first != SSG_NIL) {
        unsigned bestCount = 0;
        for (unsigned i = first; i < nRuns && (unsigned)(runKeys[i] >> valueBits) == (unsigned)s; i++) {
            const long long v = (long long)(runKeys[i] & mask) + valueMin;
            const unsigned c = runCounts[i];
            if (i == first) vmin = v;
            vmax = v;
            if (c > bestCount) { bestCount = c; mode = v; }
            pixCount += c;
            sumVC += v * (long long)c;
            last = i + 1;
        }
    }
    totalOut[s] = (unsigned)pixCount + nullCount[s];
    const bool empty = pixCount == 0;
    // (the reference keeps pixCount in a uint32 field)
    const unsigned pc = (unsigned)pixCount;
    float mean = 0.0f, stddev = 0.0f;
    if (!empty) {
        mean = (float)((double)sumVC / (double)pc);
        // tilingstats.py:960 as numba compiles it: the fused array expression is float64 per element,
        // its result array float32, and the sum of that array accumulates in float32 in index order
        float acc = 0.0f;
        for (unsigned i = first; i < last; i++) {
            const double v = (double)((long long)(runKeys[i] & mask) + valueMin);
            const double d = v - (double)mean;
            const float w = (float)((double)runCounts[i] * (d * d));
            acc = __fadd_rn(acc, w);
        }
        stddev = (float)sqrt((double)acc / (double)pc);
    }
    for (int i = 0; i < sel.n; i++) {
        long long iv = sel.statId[i] == STAT_PIXCOUNT ? 0ll : missing;   // (getStat: the field itself)
        float fv = (float)missing;
        if (!empty) {
            switch (sel.statId[i]) {
            case STAT_MIN: iv = vmin; break;
            case STAT_MAX: iv = vmax; break;
            case STAT_MEAN: fv = mean; break;
            case STAT_STDDEV: fv = stddev; break;
            case STAT_MODE: iv = mode; break;
            case STAT_PIXCOUNT: iv = (long long)pc; break;
            case STAT_MEDIAN:
            case STAT_PERCENTILE: {
                const double pct = sel.statId[i] == STAT_MEDIAN ? 50.0 : (double)(unsigned)sel.param[i];
                const double countAt = (double)pc * (pct / 100.0);
                unsigned long long cum = 0;
                unsigned j = first;
                while ((double)cum < countAt) { cum += runCounts[j]; j++; }
                // (j == first only if countAt <= 0, i.e. percentile 0: the reference then reads
                // pixVals[-1], the largest value)
                iv = (long long)(runKeys[j == first ? last - 1 : j - 1] & mask) + valueMin;
                break;
            }
            default: break;
            }
        }
        if (sel.isFloat[i]) floatOut[(size_t)sel.column[i] * len + s] = fv;
        else intOut[(size_t)sel.column[i] * len + s] = iv;
    }
}

template <typename T>
static int stats_t(ssg_ctx *ctx, const unsigned *segDev, const T *imgDev, int64_t N, int hasNull, long long nullVal,
                   uint32_t maxSegId, const StatSel &sel, long long missing, int nInt, int nFloat,
                   long long *intOutHost, float *floatOutHost, uint32_t *totalOutHost)
{
    const int64_t len = (int64_t)maxSegId + 1;
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);
    const long long valueMin = std::is_signed<T>::value ? -(1ll << (8 * sizeof(T) - 1)) : 0ll;
    const int valueBits = 8 * (int)sizeof(T);
    int segBits = 1;
    while ((1ull << segBits) <= (unsigned long long)maxSegId) segBits++;
    SSG_TRY(ssg_reserve(ctx, ctx->sortKeys0, (size_t)N * sizeof(unsigned long long)));
    SSG_TRY(ssg_reserve(ctx, ctx->sortKeys1, (size_t)N * sizeof(unsigned long long)));
    unsigned long long *keys = bufp<unsigned long long>(ctx->sortKeys0), *keys2 = bufp<unsigned long long>(ctx->sortKeys1);
    int64_t blocks = (N + 255) / 256;
    if (blocks > (int64_t)ctx->numSMs * 32) blocks = (int64_t)ctx->numSMs * 32;
    // pixels of each segment that hold the image's null value, then (k_stats_eval) all its pixels
    SSG_TRY(ssg_reserve(ctx, ctx->aux0, (size_t)len * 3 * sizeof(unsigned)));
    unsigned *segStart = bufp<unsigned>(ctx->aux0), *nullCount = segStart + len, *totalOut = nullCount + len;
    SSG_CUDA(ctx, cudaMemsetAsync(nullCount, 0, (size_t)len * sizeof(unsigned), ctx->stream));
    SSG_CUDA(ctx, cudaMemsetAsync(counters + C_SCRATCH2, 0, sizeof(unsigned long long), ctx->stream));
    SSG_PROF_BEGIN(ctx, "k_stats_keys");
    k_stats_keys<T><<<(unsigned)blocks, 256, 0, ctx->stream>>>(segDev, imgDev, N, hasNull, nullVal, valueMin, valueBits, maxSegId,
                                                               nullCount, counters + C_SCRATCH2, keys);
    SSG_LAUNCHED(ctx);
    // valid keys only
    unsigned long long *dNum = counters + C_SCRATCH0;
    size_t tmpBytes = 0;
    SSG_CUDA(ctx, cub::DeviceSelect::If(nullptr, tmpBytes, keys, keys2, dNum, N, IsValidKey(), ctx->stream));
    SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
    SSG_PROF_BEGIN(ctx, "cub_DeviceSelect_If");
    SSG_CUDA(ctx, cub::DeviceSelect::If(ctx->cubTemp.p, tmpBytes, keys, keys2, dNum, N, IsValidKey(), ctx->stream));
    SSG_LAUNCHED(ctx);
    SSG_TRY(ssg_fetch_counters(ctx));
    const int64_t M = (int64_t)ctx->hostCounters[C_SCRATCH0];
    if (ctx->hostCounters[C_SCRATCH2] != 0)
        SSG_FAIL(ctx, SSG_ERR_ARG, "%llu pixels hold a segment id above maxSegId %u",
                 (unsigned long long)ctx->hostCounters[C_SCRATCH2], maxSegId);
    int64_t nRuns = 0;
    const unsigned long long *runKeys = nullptr;
    const unsigned *runCounts = nullptr;
    SSG_CUDA(ctx, cudaMemsetAsync(segStart, 0xff, (size_t)len * sizeof(unsigned), ctx->stream));
    if (M > 0) {
        // sort by (segment, value): keys2 -> keys
        SSG_CUDA(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, keys2, keys, M, 0, valueBits + segBits, ctx->stream));
        SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
        SSG_PROF_BEGIN(ctx, "cub_DeviceRadixSort_SortKeys");
        SSG_CUDA(ctx, cub::DeviceRadixSort::SortKeys(ctx->cubTemp.p, tmpBytes, keys2, keys, M, 0, valueBits + segBits, ctx->stream));
        SSG_LAUNCHED(ctx);
        // the sorted histograms: distinct (segment, value) with their counts
        SSG_TRY(ssg_reserve(ctx, ctx->aux1, (size_t)M * sizeof(unsigned)));
        unsigned *cnts = bufp<unsigned>(ctx->aux1);
        unsigned long long *dRuns = counters + C_SCRATCH1;
        SSG_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(nullptr, tmpBytes, keys, keys2, cnts, dRuns, M, ctx->stream));
        SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
        SSG_PROF_BEGIN(ctx, "cub_DeviceRunLengthEncode_Encode");
        SSG_CUDA(ctx, cub::DeviceRunLengthEncode::Encode(ctx->cubTemp.p, tmpBytes, keys, keys2, cnts, dRuns, M, ctx->stream));
        SSG_LAUNCHED(ctx);
        SSG_TRY(ssg_fetch_counters(ctx));
        nRuns = (int64_t)ctx->hostCounters[C_SCRATCH1];
        runKeys = keys2;
        runCounts = cnts;
        SSG_PROF_BEGIN(ctx, "k_stats_starts");
        k_stats_starts<<<gridFor(nRuns, 256), 256, 0, ctx->stream>>>(runKeys, nRuns, valueBits, segStart);
        SSG_LAUNCHED(ctx);
    }
    SSG_TRY(ssg_reserve(ctx, ctx->aux2, (size_t)len * ((size_t)(nInt ? nInt : 1) * sizeof(long long) + (size_t)(nFloat ? nFloat : 1) * sizeof(float)) + 64));
    long long *intOut = bufp<long long>(ctx->aux2);
    float *floatOut = reinterpret_cast<float *>(intOut + (size_t)(nInt ? nInt : 1) * len);
    SSG_PROF_BEGIN(ctx, "k_stats_eval");
    k_stats_eval<<<gridFor(len, 128), 128, 0, ctx->stream>>>(runKeys, runCounts, nRuns, segStart, len, valueBits, valueMin, missing,
                                                           sel, nullCount, intOut, floatOut, totalOut);
    SSG_LAUNCHED(ctx);
    if (totalOutHost) SSG_CUDA(ctx, cudaMemcpyAsync(totalOutHost, totalOut, (size_t)len * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    if (nInt) SSG_CUDA(ctx, cudaMemcpyAsync(intOutHost, intOut, (size_t)nInt * len * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (nFloat) SSG_CUDA(ctx, cudaMemcpyAsync(floatOutHost, floatOut, (size_t)nFloat * len * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SSG_OK;
}

// seg / img: device pointers when onDevice, host pointers otherwise
extern "C" int ssg_segment_stats(ssg_ctx *ctx, const uint32_t *seg, const void *img, int dtype, int64_t nPixels,
                                 int onDevice, int hasNull, int64_t nullVal, uint32_t maxSegId, int nStats,
                                 const int32_t *statIds, const int32_t *params, int64_t missing,
                                 int64_t *intOut, float *floatOut, uint32_t *totalOut)
{
    if (!ctx) return SSG_ERR_ARG;
    ctx->err.clear();
    SSG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!seg || !img || nPixels < 0 || nStats < 1 || nStats > 32 || !statIds)
        SSG_FAIL(ctx, SSG_ERR_ARG, "bad argument (1..32 statistics)");
    if (dtype < SSG_U8 || dtype > SSG_I32) SSG_FAIL(ctx, SSG_ERR_ARG, "unsupported dtype code %d", dtype);
    SSG_TRY(ssg_scratch_reset(ctx));
    StatSel sel = {};
    sel.n = nStats;
    int nInt = 0, nFloat = 0;
    for (int i = 0; i < nStats; i++) {
        if (statIds[i] < STAT_MIN || statIds[i] > STAT_PIXCOUNT) SSG_FAIL(ctx, SSG_ERR_ARG, "unknown statistic id %d", statIds[i]);
        sel.statId[i] = statIds[i];
        sel.param[i] = params ? params[i] : 0;
        sel.isFloat[i] = statIds[i] == STAT_MEAN || statIds[i] == STAT_STDDEV;
        sel.column[i] = sel.isFloat[i] ? nFloat++ : nInt++;
    }
    if ((nInt && !intOut) || (nFloat && !floatOut)) SSG_FAIL(ctx, SSG_ERR_ARG, "null output pointer");
    const unsigned *segDev = seg;
    const void *imgDev = img;
    if (!onDevice) {
        const size_t ib = (size_t)nPixels * dtypeSize(dtype);
        SSG_TRY(ssg_reserve(ctx, ctx->seg, (size_t)(nPixels ? nPixels : 1) * sizeof(uint32_t)));
        SSG_TRY(ssg_reserve(ctx, ctx->img, ib ? ib : 16));
        SSG_CUDA(ctx, cudaMemcpyAsync(ctx->seg.p, seg, (size_t)nPixels * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        SSG_CUDA(ctx, cudaMemcpyAsync(ctx->img.p, img, ib, cudaMemcpyHostToDevice, ctx->stream));
        segDev = bufp<unsigned>(ctx->seg);
        imgDev = ctx->img.p;
    }
    switch (dtype) {
    case SSG_U8: return stats_t<uint8_t>(ctx, segDev, (const uint8_t *)imgDev, nPixels, hasNull, nullVal, maxSegId, sel, missing, nInt, nFloat, (long long *)intOut, floatOut, totalOut);
    case SSG_U16: return stats_t<uint16_t>(ctx, segDev, (const uint16_t *)imgDev, nPixels, hasNull, nullVal, maxSegId, sel, missing, nInt, nFloat, (long long *)intOut, floatOut, totalOut);
    case SSG_U32: return stats_t<uint32_t>(ctx, segDev, (const uint32_t *)imgDev, nPixels, hasNull, nullVal, maxSegId, sel, missing, nInt, nFloat, (long long *)intOut, floatOut, totalOut);
    case SSG_I32: return stats_t<int32_t>(ctx, segDev, (const int32_t *)imgDev, nPixels, hasNull, nullVal, maxSegId, sel, missing, nInt, nFloat, (long long *)intOut, floatOut, totalOut);
    default: return stats_t<int16_t>(ctx, segDev, (const int16_t *)imgDev, nPixels, hasNull, nullVal, maxSegId, sel, missing, nInt, nFloat, (long long *)intOut, floatOut, totalOut);
    }
}
