"""
CPU restatement of the per-segment statistics of pyshepseg.tilingstats -- TEST INFRASTRUCTURE, not
product: only tests/ may import it.  Pinned to the unmodified reference by tests/golden/stats_*.npz
(tests/golden/make_golden_stats.py).

Follows accumulateSegDict (tilingstats.py:467-517), SegmentStats (923-1008) and the column types
of RatPage (1972-1996): a histogram of the valid values per segment, sorted by value; int64
columns for min / max / median / mode / percentile / pixcount, float32 for mean / stddev.
"""
import numpy

STATIDS = {'min': 0, 'max': 1, 'mean': 2, 'stddev': 3, 'median': 4, 'mode': 5, 'percentile': 6,
    'pixcount': 7}


def calcPerSegmentStats(img, seg, statsSelection, missingStatsValue=-9999, imgNullVal=None):
    """dict column name -> array of maxSegId + 1 entries (row 0, the null segment, is zero)"""
    img = numpy.asarray(img).astype(numpy.int64).ravel()
    seg = numpy.asarray(seg).ravel()
    n = int(seg.max()) + 1 if seg.size else 1
    valid = seg != 0
    if imgNullVal is not None:
        valid &= img != int(imgNullVal)
    order = numpy.lexsort((img[valid], seg[valid]))
    s = seg[valid][order]
    v = img[valid][order]
    cols = {}
    for sel in statsSelection:
        isFloat = sel[1] in ('mean', 'stddev')
        cols[sel[0]] = numpy.zeros(n, dtype=numpy.float32 if isFloat else numpy.int64)
    bounds = numpy.searchsorted(s, numpy.arange(n + 1))
    for segId in range(1, n):
        vals = v[bounds[segId]:bounds[segId + 1]]
        if len(vals) == 0:
            # (getStat returns the pixCount field, 0, not missingStatsValue: tilingstats.py:1006)
            for sel in statsSelection:
                cols[sel[0]][segId] = 0 if sel[1] == 'pixcount' else missingStatsValue
            continue
        (pixVals, counts) = numpy.unique(vals, return_counts=True)
        counts = counts.astype(numpy.uint32)
        pixCount = numpy.uint32(counts.sum())
        mean = numpy.float32((pixVals * counts.astype(numpy.int64)).sum() / float(pixCount))
        # tilingstats.py:960 as numba compiles it: the array expression counts * (pixVals - mean)**2
        # is fused into one loop whose scalar arithmetic is float64 (int64 - float32 -> float64) and
        # whose RESULT array is float32, and .sum() of a float32 array accumulates sequentially in
        # float32; the division by the uint32 pixCount and the sqrt are float64, the field store
        # rounds to float32.
        f32 = numpy.float32
        w = (counts.astype(numpy.float64) * (pixVals.astype(numpy.float64) - float(mean)) ** 2).astype(f32)
        acc = f32(0)
        for x in w:
            acc = f32(acc + x)
        stddev = f32(numpy.sqrt(float(acc) / float(pixCount)))

        def percentile(p):
            countAt = float(pixCount) * (p / 100)
            cum = 0
            i = 0
            while cum < countAt:
                cum += int(counts[i])
                i += 1
            return pixVals[i - 1]       # (i == 0: the last value, as numba wraps the index)
        for sel in statsSelection:
            name = sel[1]
            if name == 'min':
                val = pixVals[0]
            elif name == 'max':
                val = pixVals[-1]
            elif name == 'mean':
                val = mean
            elif name == 'stddev':
                val = stddev
            elif name == 'median':
                val = percentile(50)
            elif name == 'mode':
                val = pixVals[numpy.argmax(counts)]
            elif name == 'percentile':
                val = percentile(int(sel[2]))
            elif name == 'pixcount':
                val = pixCount
            else:
                raise ValueError('unknown statistic %r' % (name,))
            cols[sel[0]][segId] = val
    return cols
