// merge.cuh -- interface between eliminate.cu (which prepares spectra, buckets and pixel lists) and
// merge.cu (the passes themselves, region mode).
#pragma once
#include "common.cuh"

struct SmallBarrier {
    unsigned arrive;
    unsigned pad[31];
};

// words of one segment record {float means[NBMAX]; uint size; uint listOffset; pad}: 32 / 64 / 128 bytes
template <int NBMAX>
struct MergeRecWords { static constexpr int value = NBMAX <= 4 ? 8 : (NBMAX <= 8 ? 16 : 32); };

struct MergeState {
    unsigned *seg;
    unsigned *segSize;                 // kept in step for the relabel that follows (0 = dead)
    unsigned *rec;                     // len records {float32 band means, size, list offset}
    float *fsum;                       // len x nB float32 band sums (touched by merges only)
    unsigned long long colMagic;       // floor(2^40 / nCols) + 1: row of a pixel index without a division
    unsigned *pix;                     // per-segment list regions of regionCap entries
    unsigned *mergeTo;                 // len, zeroed
    unsigned long long *pendHead;      // len, zeroed: (pass stamp << 32) | last pushed source
    unsigned *pendNext;                // len
    const unsigned *bucketStart;       // minSegSize + 1
    const unsigned *bucketList;        // initially small segments grouped by size
    unsigned *grownList;               // per size: targets that grew to exactly that size
    const unsigned long long *grownStart;   // minSegSize + 1 offsets into grownList
    unsigned *grownCount;              // minSegSize + 1, zeroed
    unsigned long long *ctr;           // counters (2 x MC_COUNT), then the SmallBarrier
    SmallBarrier *bar;                 // one per launch of the chain
    unsigned long long *dbg;
    unsigned safe;                     // debugging switches (SSG_MERGE_SAFE)
    unsigned stage;                    // which launch of the chain this is (0 wide grid, 1 lean grid, 2 cluster)
    unsigned long long exitSlots;      // stop at the first size >= switchMinT whose candidates x lanes fit this (0: run to the end)
    unsigned switchMinT;
    int nB;
    unsigned nRows, nCols;
    int four;
    int minSegSize;
    double thr;
};

struct MergePlan {
    const float *fsum;          // len x nB float32 band sums (buildSegmentSpectra)
    const unsigned *sliceOff;   // len: list region offset of every segment
    int64_t len;
};

size_t ssgk_merge_rec_bytes(int nB, int64_t len);
size_t ssgk_merge_ctr_bytes(void);
int ssgk_merge_regions(ssg_ctx *ctx, MergeState &st, const MergePlan &plan, uint32_t *numPasses, int64_t *numElim);
