// group.cu -- "pixels of flagged segments, grouped by segment, raster order inside".
// The one place the library leans on CUB's select and radix sort.  It serves only the rare
// exact-order paths: replaying the clump cap on oversized regions (shepseg.py:481-539) and
// the ordered float32 accumulation of segment sums that are not exactly representable
// (shepseg.py:805-811).
#include "common.cuh"

#include <cub/device/device_select.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <thrust/iterator/counting_iterator.h>

struct InFlaggedSegment {
    const unsigned *seg;
    const unsigned char *flag;
    __device__ bool operator()(unsigned p) const { return flag[seg[p]] != 0; }
};

struct IsRunStart {
    const unsigned *keys;
    __device__ bool operator()(unsigned i) const { return i == 0 || keys[i] != keys[i - 1]; }
};

__global__ void __launch_bounds__(256)
k_fetch_keys(const unsigned *__restrict__ pix, int64_t M, const unsigned *__restrict__ seg,
             unsigned *keys)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) keys[i] = seg[pix[i]];
}

int ssgk_group_pixels(ssg_ctx *ctx, const unsigned *segDev, int64_t N, const unsigned char *segFlag,
                      const unsigned **pixSortedOut, const unsigned **keysSortedOut,
                      const unsigned **runStartOut, int64_t *MOut, unsigned *numRunsOut)
{
    unsigned long long *counters = bufp<unsigned long long>(ctx->counters);
    *MOut = 0;
    *numRunsOut = 0;
    if (N >= 0x7FFFFFFFll) SSG_FAIL(ctx, SSG_ERR_ARG, "tile too large for the grouped-pixel path");
    SSG_TRY(ssg_reserve(ctx, ctx->sortVals0, (size_t)N * sizeof(unsigned)));
    unsigned *pix = bufp<unsigned>(ctx->sortVals0);
    InFlaggedSegment pred{segDev, segFlag};
    thrust::counting_iterator<unsigned> cnt(0);
    unsigned long long *dNum = counters + C_SCRATCH0;
    size_t tmpBytes = 0;
    SSG_CUDA(ctx, cub::DeviceSelect::If(nullptr, tmpBytes, cnt, pix, dNum, (int)N, pred, ctx->stream));
    SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
    SSG_PROF_BEGIN(ctx, "cub_DeviceSelect_If");
    SSG_CUDA(ctx, cub::DeviceSelect::If(ctx->cubTemp.p, tmpBytes, cnt, pix, dNum, (int)N, pred, ctx->stream));
    SSG_LAUNCHED(ctx);
    SSG_TRY(ssg_fetch_counters(ctx));
    const int64_t M = (int64_t)(ctx->hostCounters[C_SCRATCH0] & 0xffffffffu);
    if (M == 0) return SSG_OK;

    SSG_TRY(ssg_reserve(ctx, ctx->sortKeys0, (size_t)M * sizeof(unsigned)));
    SSG_TRY(ssg_reserve(ctx, ctx->sortKeys1, (size_t)M * sizeof(unsigned)));
    SSG_TRY(ssg_reserve(ctx, ctx->sortVals1, (size_t)M * sizeof(unsigned)));
    unsigned *keys0 = bufp<unsigned>(ctx->sortKeys0), *keys1 = bufp<unsigned>(ctx->sortKeys1);
    unsigned *pixSorted = bufp<unsigned>(ctx->sortVals1);
    SSG_PROF_BEGIN(ctx, "k_fetch_keys");
    k_fetch_keys<<<gridFor(M, 256), 256, 0, ctx->stream>>>(pix, M, segDev, keys0);
    SSG_LAUNCHED(ctx);
    // LSD radix sort is stable: raster order survives inside each segment
    SSG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys0, keys1, pix, pixSorted, (int)M, 0, 32, ctx->stream));
    SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
    SSG_PROF_BEGIN(ctx, "cub_DeviceRadixSort_SortPairs");
    SSG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->cubTemp.p, tmpBytes, keys0, keys1, pix, pixSorted, (int)M, 0, 32, ctx->stream));
    SSG_LAUNCHED(ctx);
    unsigned *runStart = keys0;   // keys0 is free again; at most M runs
    IsRunStart rs{keys1};
    SSG_CUDA(ctx, cub::DeviceSelect::If(nullptr, tmpBytes, cnt, runStart, dNum, (int)M, rs, ctx->stream));
    SSG_TRY(ssg_reserve(ctx, ctx->cubTemp, tmpBytes));
    SSG_PROF_BEGIN(ctx, "cub_DeviceSelect_If");
    SSG_CUDA(ctx, cub::DeviceSelect::If(ctx->cubTemp.p, tmpBytes, cnt, runStart, dNum, (int)M, rs, ctx->stream));
    SSG_LAUNCHED(ctx);
    SSG_TRY(ssg_fetch_counters(ctx));
    *numRunsOut = (unsigned)(ctx->hostCounters[C_SCRATCH0] & 0xffffffffu);
    *MOut = M;
    *pixSortedOut = pixSorted;
    if (keysSortedOut) *keysSortedOut = keys1;
    *runStartOut = runStart;
    return SSG_OK;
}
