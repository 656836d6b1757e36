/*
 * shepseg_oracle.c -- CPU restatement of pyshepseg's Shepherd-segmentation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * pyshepseg_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here
 * against label rasters produced by the unmodified reference (pyshepseg 2.0.3 run
 * under numba 0.65.0 / scikit-learn 1.9.0 / scipy 1.18.1 / numpy 2.3.5) and stored
 * under tests/golden/ by tests/golden/make_golden.py.
 *
 * Plain C, single threaded like the reference's numba code, except orc_assign which
 * uses OpenMP like scikit-learn's predict does.  Compile with -ffp-contract=off so
 * no FMA contraction changes the float32/float64 sequences restated here.
 *
 * Each function cites the reference lines (pyshepseg/shepseg.py, pyshepseg/tiling.py)
 * whose behaviour it restates.  Semantics that are not visible in the Python source
 * (they come from numba's typing rules) were pinned by inspecting the LLVM IR numba
 * generates:
 *   - spectSum[s,k] += img[k,i,j] is RN32(exact(s + x))            (shepseg.py:811)
 *   - spectSum[s] / n divides in float64 and truncates to float32  (shepseg.py:1042,1053)
 *   - ((a-b)**2).sum() on float32 arrays is a sequential float32 chain, no FMA (1055)
 *   - img[:,i,j]-img[:,ii,jj] widens to int64 before subtracting   (shepseg.py:730)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_U8 0
#define ORC_U16 1
#define ORC_I16 2
#define ORC_U32 3
#define ORC_I32 4

#define ORC_MAX_BANDS 64
#define ORC_MAX_CLUMP_SIZE 10000 /* shepseg.py:481 */

typedef uint32_t segid_t; /* shepseg.py:97 SegIdType */

static inline int64_t px_get(const void *img, int dt, size_t idx)
{
    switch (dt) {
    case ORC_U8:  return ((const uint8_t *)img)[idx];
    case ORC_U16: return ((const uint16_t *)img)[idx];
    case ORC_I16: return ((const int16_t *)img)[idx];
    case ORC_U32: return ((const uint32_t *)img)[idx];
    default:      return ((const int32_t *)img)[idx];
    }
}

/* ------------------------------------------------------------------------------------
 * applySpectralClusters  (shepseg.py:317-361)
 *
 * kmeansObj.predict (scikit-learn 1.9.0, sklearn/cluster/_k_means_lloyd.pyx
 * _update_chunk_dense, NOT part of /root/reference) converts the pixels to float64 and,
 * per sample, evaluates  ||c_j||^2 - 2 * x . c_j  for every centre j and keeps the first
 * strict minimum.  Cluster numbers are made 1-based (shepseg.py:356) and pixels with any
 * band equal to imgNullVal become 0 (shepseg.py:357-359).
 * ---------------------------------------------------------------------------------- */
int orc_assign(const void *img, int dt, int nBands, int64_t nRows, int64_t nCols,
               const double *centres, int k, int hasNull, double nullVal, int32_t *out)
{
    if (nBands > ORC_MAX_BANDS || nBands < 1 || k < 1) return 1;
    const size_t N = (size_t)nRows * (size_t)nCols;
    double *cn = (double *)malloc(sizeof(double) * (size_t)k);
    if (!cn) return 2;
    for (int j = 0; j < k; j++) {
        double s = 0.0;
        for (int b = 0; b < nBands; b++) s += centres[j * nBands + b] * centres[j * nBands + b];
        cn[j] = s;
    }
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)N; p++) {
        double x[ORC_MAX_BANDS];
        int isNull = 0;
        for (int b = 0; b < nBands; b++) {
            x[b] = (double)px_get(img, dt, (size_t)b * N + (size_t)p);
            if (hasNull && x[b] == nullVal) isNull = 1;
        }
        if (isNull) { out[p] = 0; continue; }
        int best = 0;
        double bestD = 0.0;
        for (int j = 0; j < k; j++) {
            double dot = 0.0;
            for (int b = 0; b < nBands; b++) dot += x[b] * centres[j * nBands + b];
            double d = cn[j] + (-2.0 * dot);
            if (j == 0 || d < bestD) { bestD = d; best = j; }
        }
        out[p] = best + 1;
    }
    free(cn);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * clump  (shepseg.py:452-541)
 * Raster scan; each unvisited non-ignore pixel seeds a LIFO flood fill over equal
 * values.  The fill stops popping once MAX_CLUMP_SIZE pixels have been pushed
 * (shepseg.py:502); neighbours are visited columns outer, rows inner (523-524).
 * ---------------------------------------------------------------------------------- */
int orc_clump(const int32_t *img, int64_t nRows, int64_t nCols, int32_t ignoreVal,
              int fourConnected, uint32_t clumpId, segid_t *out, uint32_t *nextId)
{
    const size_t N = (size_t)nRows * (size_t)nCols;
    memset(out, 0, N * sizeof(segid_t));
    uint32_t *stack = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (N ? N : 1));
    if (!stack) return 2;
    for (int64_t y = 0; y < nRows; y++) {
        for (int64_t x = 0; x < nCols; x++) {
            size_t p = (size_t)y * nCols + x;
            if (img[p] == ignoreVal || out[p] != 0) continue;
            int32_t val = img[p];
            int64_t clumpSize = 0;
            size_t sp = 0;
            stack[0] = (uint32_t)y; stack[1] = (uint32_t)x; sp = 1;
            out[p] = clumpId;
            while (sp > 0 && clumpSize < ORC_MAX_CLUMP_SIZE) {
                sp--;
                int64_t sy = stack[2 * sp], sx = stack[2 * sp + 1];
                int64_t tlx = sx - 1 < 0 ? 0 : sx - 1;
                int64_t tly = sy - 1 < 0 ? 0 : sy - 1;
                int64_t brx = sx + 1 > nCols - 1 ? nCols - 1 : sx + 1;
                int64_t bry = sy + 1 > nRows - 1 ? nRows - 1 : sy + 1;
                for (int64_t cx = tlx; cx <= brx; cx++) {
                    for (int64_t cy = tly; cy <= bry; cy++) {
                        int connected = !fourConnected || (cy == sy || cx == sx);
                        size_t q = (size_t)cy * nCols + cx;
                        if (connected && img[q] != ignoreVal && out[q] == 0 && img[q] == val) {
                            out[q] = clumpId;
                            clumpSize++;
                            stack[2 * sp] = (uint32_t)cy; stack[2 * sp + 1] = (uint32_t)cx;
                            sp++;
                        }
                    }
                }
            }
            clumpId++;
        }
    }
    free(stack);
    *nextId = clumpId;
    return 0;
}

/* makeSegSize (shepseg.py:544-569): histogram of ids; segSize has maxId+1 entries. */
int orc_make_seg_size(const segid_t *seg, int64_t N, uint32_t *segSize, int64_t len)
{
    memset(segSize, 0, sizeof(uint32_t) * (size_t)len);
    for (int64_t p = 0; p < N; p++) {
        if ((int64_t)seg[p] >= len) return 3;
        segSize[seg[p]]++;
    }
    return 0;
}

/* relabelSegments (shepseg.py:739-777): order preserving compaction of ids >= minSegId. */
int orc_relabel_segments(segid_t *seg, int64_t N, const uint32_t *segSize, int64_t len,
                         uint32_t minSegId)
{
    uint32_t *subtract = (uint32_t *)calloc((size_t)(len > 0 ? len : 1), sizeof(uint32_t));
    if (!subtract) return 2;
    for (int64_t kk = (int64_t)minSegId + 1; kk < len; kk++) {
        subtract[kk] = subtract[kk - 1];
        if (segSize[kk - 1] == 0) subtract[kk]++;
    }
    for (int64_t p = 0; p < N; p++) seg[p] = seg[p] - subtract[seg[p]];
    free(subtract);
    return 0;
}

/* findNearestNeighbourPixel (shepseg.py:677-736).  Returns linear index of the chosen
 * neighbour or -1.  Rows outer, columns inner; int64 squared distance; first strict
 * minimum wins; only neighbours whose segment has more than one pixel qualify (the null
 * segment 0 qualifies like any other). */
static int64_t nearest_neighbour_pixel(const void *img, int dt, int nBands, int64_t nRows,
                                       int64_t nCols, const segid_t *seg, int64_t i, int64_t j,
                                       const uint32_t *segSize, int fourConnected)
{
    const size_t N = (size_t)nRows * nCols;
    int64_t minD = -1, best = -1;
    int64_t i0 = i - 1 < 0 ? 0 : i - 1, i1 = i + 1 > nRows - 1 ? nRows - 1 : i + 1;
    int64_t j0 = j - 1 < 0 ? 0 : j - 1, j1 = j + 1 > nCols - 1 ? nCols - 1 : j + 1;
    for (int64_t ii = i0; ii <= i1; ii++) {
        for (int64_t jj = j0; jj <= j1; jj++) {
            int connected = !fourConnected || (ii == i || jj == j);
            if (!connected) continue;
            size_t q = (size_t)ii * nCols + jj;
            if (segSize[seg[q]] > 1) {
                uint64_t d = 0; /* wraps like numba's int64 */
                for (int b = 0; b < nBands; b++) {
                    int64_t df = px_get(img, dt, (size_t)b * N + (size_t)i * nCols + j) -
                                 px_get(img, dt, (size_t)b * N + q);
                    d += (uint64_t)df * (uint64_t)df;
                }
                int64_t ds = (int64_t)d;
                if (minD < 0 || ds < minD) { minD = ds; best = (int64_t)q; }
            }
        }
    }
    return best;
}

/* mergeSinglePixels (shepseg.py:618-674): scan phase with frozen state, then apply. */
static int64_t merge_single_pixels(const void *img, int dt, int nBands, int64_t nRows,
                                   int64_t nCols, segid_t *seg, uint32_t *segSize,
                                   uint32_t *elimPix, segid_t *elimSeg, int fourConnected)
{
    int64_t n = 0;
    for (int64_t i = 0; i < nRows; i++) {
        for (int64_t j = 0; j < nCols; j++) {
            size_t p = (size_t)i * nCols + j;
            if (segSize[seg[p]] == 1) {
                int64_t q = nearest_neighbour_pixel(img, dt, nBands, nRows, nCols, seg, i, j,
                                                    segSize, fourConnected);
                if (q >= 0) { elimPix[n] = (uint32_t)p; elimSeg[n] = seg[q]; n++; }
            }
        }
    }
    for (int64_t t = 0; t < n; t++) {
        segid_t oldSeg = seg[elimPix[t]];
        seg[elimPix[t]] = elimSeg[t];
        segSize[oldSeg] = 0;
        segSize[elimSeg[t]] += 1;
    }
    return n;
}

/* eliminateSinglePixels (shepseg.py:572-615): rounds until nothing moves, then relabel.
 * segSize (len entries) is updated in place exactly as the reference leaves it (stale
 * after the relabel).  Returns the total number of moves through *numMoved. */
int orc_eliminate_single_pixels(const void *img, int dt, int nBands, int64_t nRows,
                                int64_t nCols, segid_t *seg, uint32_t *segSize, int64_t len,
                                uint32_t minSegId, int fourConnected, int64_t *numMoved)
{
    const size_t N = (size_t)nRows * nCols;
    uint32_t *elimPix = (uint32_t *)malloc(sizeof(uint32_t) * (N ? N : 1));
    segid_t *elimSeg = (segid_t *)malloc(sizeof(segid_t) * (N ? N : 1));
    if (!elimPix || !elimSeg) { free(elimPix); free(elimSeg); return 2; }
    int64_t total = 0, n;
    do {
        n = merge_single_pixels(img, dt, nBands, nRows, nCols, seg, segSize, elimPix, elimSeg,
                                fourConnected);
        total += n;
    } while (n > 0);
    free(elimPix); free(elimSeg);
    if (numMoved) *numMoved = total;
    return orc_relabel_segments(seg, (int64_t)N, segSize, len, minSegId);
}

/* buildSegmentSpectra (shepseg.py:780-813): float32 running sums in raster order. */
static void build_segment_spectra(const void *img, int dt, int nBands, size_t N,
                                  const segid_t *seg, float *spectSum)
{
    for (size_t p = 0; p < N; p++) {
        float *row = spectSum + (size_t)seg[p] * nBands;
        for (int b = 0; b < nBands; b++)
            row[b] = (float)((double)row[b] + (double)px_get(img, dt, (size_t)b * N + p));
    }
}

#define ORC_NIL 0xFFFFFFFFu

typedef struct {
    int nBands; int64_t nRows, nCols; int fourConnected;
    segid_t *seg; uint32_t *segSize; float *spectSum;
    uint32_t *next, *head, *tail; /* per-segment pixel lists, in the reference's list order */
    double thr;                  /* maxSpectralDiff**2 evaluated in its own dtype, as double */
} small_state;

/* findMergeSegment (shepseg.py:1003-1063). */
static segid_t find_merge_segment(const small_state *st, segid_t segId)
{
    const int nB = st->nBands;
    float spect[ORC_MAX_BANDS], nbrSpect[ORC_MAX_BANDS];
    const uint32_t numPix = st->segSize[segId];
    for (int b = 0; b < nB; b++)
        spect[b] = (float)((double)st->spectSum[(size_t)segId * nB + b] / (double)numPix);
    segid_t best = 0;
    double bestD = 0.0;
    for (uint32_t p = st->head[segId]; p != ORC_NIL; p = st->next[p]) {
        int64_t i = p / st->nCols, j = p % st->nCols;
        int64_t i0 = i - 1 < 0 ? 0 : i - 1, i1 = i + 2 > st->nRows ? st->nRows : i + 2;
        int64_t j0 = j - 1 < 0 ? 0 : j - 1, j1 = j + 2 > st->nCols ? st->nCols : j + 2;
        for (int64_t ii = i0; ii < i1; ii++) {
            for (int64_t jj = j0; jj < j1; jj++) {
                int connected = !st->fourConnected || (ii == i || jj == j);
                segid_t nbr = st->seg[(size_t)ii * st->nCols + jj];
                if (connected && nbr != segId && nbr != 0 && st->segSize[nbr] > st->segSize[segId]) {
                    const uint32_t nsz = st->segSize[nbr];
                    for (int b = 0; b < nB; b++)
                        nbrSpect[b] = (float)((double)st->spectSum[(size_t)nbr * nB + b] / (double)nsz);
                    float d = 0.0f;
                    for (int b = 0; b < nB; b++) {
                        float df = spect[b] - nbrSpect[b];
                        float sq = df * df;
                        d = d + sq;
                    }
                    if (best == 0 || (double)d < bestD) { bestD = (double)d; best = nbr; }
                }
            }
        }
    }
    if (bestD > st->thr) best = 0;
    return best;
}

/* doMerge (shepseg.py:1066-1123): relabel, list = target's list ++ source's list,
 * float32 sum add, size add. */
static void do_merge(small_state *st, segid_t segId, segid_t nbr)
{
    for (uint32_t p = st->head[segId]; p != ORC_NIL; p = st->next[p]) st->seg[p] = nbr;
    if (st->head[segId] != ORC_NIL) {
        if (st->head[nbr] == ORC_NIL) st->head[nbr] = st->head[segId];
        else st->next[st->tail[nbr]] = st->head[segId];
        st->tail[nbr] = st->tail[segId];
    }
    st->head[segId] = st->tail[segId] = ORC_NIL;
    for (int b = 0; b < st->nBands; b++) {
        float *t = &st->spectSum[(size_t)nbr * st->nBands + b];
        float *s = &st->spectSum[(size_t)segId * st->nBands + b];
        *t = *t + *s;
        *s = 0.0f;
    }
    st->segSize[nbr] += st->segSize[segId];
    st->segSize[segId] = 0;
}

/* eliminateSmallSegments (shepseg.py:918-1000).  thr is maxSpectralDiff**2 as the
 * reference evaluates it (float32 product for numpy.float32, float64 for a Python float),
 * widened to double by the caller. */
int orc_eliminate_small_segments(segid_t *seg, const void *img, int dt, int nBands,
                                 int64_t nRows, int64_t nCols, uint32_t maxSegId,
                                 int minSegSize, double thr, int fourConnected,
                                 uint32_t minSegId, int64_t *numElimOut)
{
    if (nBands > ORC_MAX_BANDS) return 1;
    const size_t N = (size_t)nRows * nCols;
    const size_t S = (size_t)maxSegId + 1;
    small_state st;
    st.nBands = nBands; st.nRows = nRows; st.nCols = nCols; st.fourConnected = fourConnected;
    st.seg = seg; st.thr = thr;
    st.spectSum = (float *)calloc(S * nBands, sizeof(float));
    st.segSize = (uint32_t *)calloc(S, sizeof(uint32_t));
    st.next = (uint32_t *)malloc(sizeof(uint32_t) * (N ? N : 1));
    st.head = (uint32_t *)malloc(sizeof(uint32_t) * S);
    st.tail = (uint32_t *)malloc(sizeof(uint32_t) * S);
    segid_t *mergeSeg = (segid_t *)calloc(S, sizeof(segid_t));
    if (!st.spectSum || !st.segSize || !st.next || !st.head || !st.tail || !mergeSeg) return 2;

    build_segment_spectra(img, dt, nBands, N, seg, st.spectSum);
    for (size_t p = 0; p < N; p++) st.segSize[seg[p]]++;
    /* makeSegmentLocations (shepseg.py:880-915): raster-order pixel list per segment */
    for (size_t s = 0; s < S; s++) st.head[s] = st.tail[s] = ORC_NIL;
    for (size_t p = 0; p < N; p++) {
        segid_t s = seg[p];
        st.next[p] = ORC_NIL;
        if (s == 0) continue;
        if (st.head[s] == ORC_NIL) st.head[s] = (uint32_t)p; else st.next[st.tail[s]] = (uint32_t)p;
        st.tail[s] = (uint32_t)p;
    }

    int64_t numElim = 0;
    for (int targetSize = 1; targetSize < minSegSize; targetSize++) {
        int64_t count = 0, prev = -1;
        for (size_t s = 0; s < S; s++) count += (st.segSize[s] == (uint32_t)targetSize);
        int numPasses = 0;
        while (count != prev && numPasses < 10) {
            prev = count;
            for (size_t s = minSegId; s < S; s++)
                if (st.segSize[s] == (uint32_t)targetSize)
                    mergeSeg[s] = find_merge_segment(&st, (segid_t)s);
            for (size_t s = minSegId; s < S; s++) {
                if (mergeSeg[s] != 0) {
                    do_merge(&st, (segid_t)s, mergeSeg[s]);
                    mergeSeg[s] = 0;
                    numElim++;
                }
            }
            count = 0;
            for (size_t s = 0; s < S; s++) count += (st.segSize[s] == (uint32_t)targetSize);
            numPasses++;
        }
    }
    int rc = orc_relabel_segments(seg, (int64_t)N, st.segSize, (int64_t)S, minSegId);
    free(st.spectSum); free(st.segSize); free(st.next); free(st.head); free(st.tail); free(mergeSeg);
    if (numElimOut) *numElimOut = numElim;
    return rc;
}

/* doShepherdSegmentation (shepseg.py:130-249) with kmeansObj given: assign, clump,
 * sizes, single pixels, small segments.  thr as for orc_eliminate_small_segments. */
int orc_segment(const void *img, int dt, int nBands, int64_t nRows, int64_t nCols,
                const double *centres, int k, int hasNull, double nullVal, int fourConnected,
                int minSegSize, double thr, segid_t *segOut, uint32_t *numSegments,
                uint32_t *numElimSingle, int64_t *numElimSmall, uint32_t *numClumps)
{
    const size_t N = (size_t)nRows * nCols;
    int32_t *clusters = (int32_t *)malloc(sizeof(int32_t) * (N ? N : 1));
    if (!clusters) return 2;
    int rc = orc_assign(img, dt, nBands, nRows, nCols, centres, k, hasNull, nullVal, clusters);
    if (rc) { free(clusters); return rc; }
    uint32_t nextId = 0;
    rc = orc_clump(clusters, nRows, nCols, 0, fourConnected, 1, segOut, &nextId);
    free(clusters);
    if (rc) return rc;
    uint32_t maxSegId = nextId - 1; /* shepseg.py:214 */
    if (numClumps) *numClumps = maxSegId;
    uint32_t *segSize = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)maxSegId + 1));
    if (!segSize) return 2;
    orc_make_seg_size(segOut, (int64_t)N, segSize, (int64_t)maxSegId + 1);
    rc = orc_eliminate_single_pixels(img, dt, nBands, nRows, nCols, segOut, segSize,
                                     (int64_t)maxSegId + 1, 1, fourConnected, NULL);
    free(segSize);
    if (rc) return rc;
    uint32_t newMax = 0;
    for (size_t p = 0; p < N; p++) if (segOut[p] > newMax) newMax = segOut[p];
    if (numElimSingle) *numElimSingle = maxSegId - newMax; /* shepseg.py:226-227 */
    int64_t ne = 0;
    rc = orc_eliminate_small_segments(segOut, img, dt, nBands, nRows, nCols, newMax, minSegSize,
                                      thr, fourConnected, 1, &ne);
    if (rc) return rc;
    if (numElimSmall) *numElimSmall = ne;
    uint32_t finalMax = 0;
    for (size_t p = 0; p < N; p++) if (segOut[p] > finalMax) finalMax = segOut[p];
    if (numSegments) *numSegments = finalMax;
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Tile stitching: recodeTile (tiling.py:1066-1126), recodeSharedSegments (1128-1203),
 * crossesMidline (1271-1306), relabelSegments (1205-1269).
 *
 * tile is (ysize, xsize); topB is the upper neighbour's recoded bottom strip
 * (overlap, xsize) or NULL; leftB is the left neighbour's recoded right strip
 * (ysize, overlap) or NULL.  Tile ids must be contiguous 1..max (true for
 * doShepherdSegmentation output).  out receives the recoded copy, *newMaxSegId the
 * running id counter after numbering this tile's own segments.
 * ---------------------------------------------------------------------------------- */
typedef struct { segid_t a, b; } pair_t;
static int pair_cmp(const void *x, const void *y)
{
    const pair_t *p = (const pair_t *)x, *q = (const pair_t *)y;
    if (p->a != q->a) return p->a < q->a ? -1 : 1;
    if (p->b != q->b) return p->b < q->b ? -1 : 1;
    return 0;
}

/* One overlap strip.  alongRows=1: strip is the top `ov` rows, stitch axis = rows
 * (HORIZONTAL, tiling.py:1296-1298); else the left `ov` columns (VERTICAL, 1299-1301).
 * recode[id] gets the mode of B over the strip pixels of id (smallest value among
 * ties, scipy.stats.mode; 0 is a legal value) and isKey[id]=1. */
static int recode_shared(const segid_t *tile, int64_t ysize, int64_t xsize, int64_t ov,
                         const segid_t *B, int alongRows, uint32_t maxId, segid_t *recode,
                         uint8_t *isKey)
{
    int64_t sRows = alongRows ? (ov < ysize ? ov : ysize) : ysize;
    int64_t sCols = alongRows ? xsize : (ov < xsize ? ov : xsize);
    int64_t bStride = sCols; /* B has the same shape as the strip */
    int64_t mid = (alongRows ? sRows : sCols) / 2;
    int64_t *mn = (int64_t *)malloc(sizeof(int64_t) * ((size_t)maxId + 1));
    int64_t *mx = (int64_t *)malloc(sizeof(int64_t) * ((size_t)maxId + 1));
    if (!mn || !mx) { free(mn); free(mx); return 2; }
    for (uint32_t s = 0; s <= maxId; s++) { mn[s] = INT64_MAX; mx[s] = -1; }
    for (int64_t r = 0; r < sRows; r++)
        for (int64_t c = 0; c < sCols; c++) {
            segid_t s = tile[r * xsize + c];
            int64_t v = alongRows ? r : c;
            if (v < mn[s]) mn[s] = v;
            if (v > mx[s]) mx[s] = v;
        }
    size_t np = 0;
    for (int64_t r = 0; r < sRows; r++)
        for (int64_t c = 0; c < sCols; c++) {
            segid_t s = tile[r * xsize + c];
            if (s != 0 && mn[s] < mid && mx[s] >= mid) np++;
        }
    pair_t *pairs = (pair_t *)malloc(sizeof(pair_t) * (np ? np : 1));
    if (!pairs) { free(mn); free(mx); return 2; }
    np = 0;
    for (int64_t r = 0; r < sRows; r++)
        for (int64_t c = 0; c < sCols; c++) {
            segid_t s = tile[r * xsize + c];
            if (s != 0 && mn[s] < mid && mx[s] >= mid) {
                pairs[np].a = s; pairs[np].b = B[r * bStride + c]; np++;
            }
        }
    qsort(pairs, np, sizeof(pair_t), pair_cmp);
    size_t i = 0;
    while (i < np) {
        segid_t s = pairs[i].a, bestVal = 0;
        size_t bestCnt = 0;
        while (i < np && pairs[i].a == s) {
            segid_t v = pairs[i].b; size_t c = 0;
            while (i < np && pairs[i].a == s && pairs[i].b == v) { c++; i++; }
            if (c > bestCnt) { bestCnt = c; bestVal = v; } /* ascending v: first max = smallest */
        }
        recode[s] = bestVal; isKey[s] = 1;
    }
    free(pairs); free(mn); free(mx);
    return 0;
}

int orc_recode_tile(const segid_t *tile, int64_t ysize, int64_t xsize, int64_t overlap,
                    const segid_t *topB, const segid_t *leftB, uint32_t maxSegId,
                    int64_t top, int64_t bottom, int64_t left, int64_t right,
                    segid_t *out, uint32_t *newMaxSegId)
{
    const size_t N = (size_t)ysize * xsize;
    uint32_t maxId = 0;
    for (size_t p = 0; p < N; p++) if (tile[p] > maxId) maxId = tile[p];
    segid_t *recode = (segid_t *)calloc((size_t)maxId + 1, sizeof(segid_t));
    uint8_t *isKey = (uint8_t *)calloc((size_t)maxId + 1, 1);
    int64_t *segTop = (int64_t *)malloc(sizeof(int64_t) * ((size_t)maxId + 1));
    int64_t *segLeft = (int64_t *)malloc(sizeof(int64_t) * ((size_t)maxId + 1));
    segid_t *lut = (segid_t *)calloc((size_t)maxId + 1, sizeof(segid_t));
    if (!recode || !isKey || !segTop || !segLeft || !lut) return 2;
    int rc = 0;
    if (topB) rc = recode_shared(tile, ysize, xsize, overlap, topB, 1, maxId, recode, isKey);
    if (!rc && leftB) rc = recode_shared(tile, ysize, xsize, overlap, leftB, 0, maxId, recode, isKey);
    if (rc) return rc;
    for (uint32_t s = 0; s <= maxId; s++) { segTop[s] = INT64_MAX; segLeft[s] = INT64_MAX; }
    for (int64_t r = 0; r < ysize; r++)
        for (int64_t c = 0; c < xsize; c++) {
            segid_t s = tile[r * xsize + c];
            if (r < segTop[s]) segTop[s] = r;
            if (c < segLeft[s]) segLeft[s] = c;
        }
    uint32_t newSegId = maxSegId;
    for (uint32_t s = 1; s <= maxId; s++) {
        if (segTop[s] == INT64_MAX) continue; /* id with no pixels: nothing to write */
        if (isKey[s]) lut[s] = recode[s];
        else if (segLeft[s] >= left && segTop[s] >= top && segLeft[s] < right && segTop[s] < bottom)
            lut[s] = ++newSegId;
        else lut[s] = 0;
    }
    for (size_t p = 0; p < N; p++) out[p] = lut[tile[p]];
    free(recode); free(isKey); free(segTop); free(segLeft); free(lut);
    if (newMaxSegId) *newMaxSegId = newSegId;
    return 0;
}
