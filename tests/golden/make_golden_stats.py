"""
Golden fixtures for the per-segment statistics: the UNMODIFIED reference
(/root/reference/pyshepseg/tilingstats.py) run through the numpy stand-in for GDAL
(tests/golden/fake_gdal.py) on small seeded rasters; inputs and the RAT columns it wrote are
stored in tests/golden/stats_*.npz.

    python tests/golden/make_golden_stats.py        (in the container that has /root/reference)
"""
import json
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fake_gdal  # noqa: E402
gdal = fake_gdal.install()


class _SR(object):
    def __init__(self, wkt=''):
        self.wkt = wkt

    def IsSame(self, other):
        return self.wkt == other.wkt


sys.modules['osgeo.osr'].SpatialReference = _SR
sys.modules['osgeo.osr'].UseExceptions = lambda: None
sys.path.insert(0, '/root/reference')
from pyshepseg import tilingstats  # noqa: E402
from pyshepseg_b200 import synth  # noqa: E402

SELECTION = [('b_min', 'min'), ('b_max', 'max'), ('b_mean', 'mean'), ('b_std', 'stddev'),
    ('b_median', 'median'), ('b_mode', 'mode'), ('b_p25', 'percentile', 25), ('b_p75', 'percentile', 75),
    ('b_p0', 'percentile', 0), ('b_p100', 'percentile', 100), ('b_count', 'pixcount')]


def labels(nR, nC, cell, seed, nullFrac):
    """irregular blobs: Voronoi cells of random seeds, ids 1..n contiguous, an optional null wedge"""
    rng = numpy.random.default_rng(seed)
    n = (nR // cell) * (nC // cell)
    cy = rng.uniform(0, nR, n)
    cx = rng.uniform(0, nC, n)
    rr = numpy.arange(nR)[:, None]
    cc = numpy.arange(nC)[None, :]
    best = numpy.full((nR, nC), numpy.inf)
    seg = numpy.zeros((nR, nC), dtype=numpy.uint32)
    for i in range(n):
        d = (rr - cy[i]) ** 2 + (cc - cx[i]) ** 2
        m = d < best
        best[m] = d[m]
        seg[m] = i + 1
    if nullFrac > 0:
        seg[(rr + cc) < numpy.sqrt(2 * nullFrac * nR * nC)] = 0
    # contiguous ids
    (u, inv) = numpy.unique(seg, return_inverse=True)
    lut = numpy.arange(len(u), dtype=numpy.uint32)
    if u[0] != 0:
        lut += 1
    return lut[inv].reshape(nR, nC).astype(numpy.uint32)


def run(name, nR, nC, dtype, cell, seed, nullFrac, imgNull):
    seg = labels(nR, nC, cell, seed, nullFrac)
    rng = numpy.random.default_rng(seed + 1)
    info = numpy.iinfo(dtype)
    base = synth.synth_v1(nR, nC, 1, seed=seed)[0].astype(numpy.int64)
    if dtype == numpy.uint8:
        img = (base // 16).clip(0, 255).astype(dtype)
    elif dtype == numpy.int16:
        img = (base - 2500).clip(info.min, info.max).astype(dtype)
    elif dtype == numpy.uint32:
        # values far above 2**24: the float32 roundings of the reference's stddev matter here
        img = ((base * 600011 + rng.integers(0, 1000, base.shape)) % 2**32).astype(dtype)
    elif dtype == numpy.int32:
        img = (base * 400009 - 2**31 + rng.integers(0, 1000, base.shape)).clip(info.min, info.max).astype(dtype)
    else:
        img = base.astype(dtype)
    if imgNull is not None:
        img[rng.random((nR, nC)) < 0.07] = imgNull
        # one segment entirely nodata
        img[seg == seg.max() // 2] = imgNull
    segSize = numpy.bincount(seg.ravel(), minlength=int(seg.max()) + 1).astype(numpy.float64)
    segSize[0] = 0
    fake_gdal.put_image('img_' + name, img[None], nodata=imgNull)
    segds = fake_gdal.put_image('seg_' + name, seg[None].copy())
    rat = segds.GetRasterBand(1).GetDefaultRAT()
    rat.SetRowCount(len(segSize))
    rat.CreateColumn('Histogram', gdal.GFT_Real, gdal.GFU_PixelCount)
    rat.WriteArray(segSize, 0)
    tilingstats.calcPerSegmentStatsTiled('img_' + name, 1, 'seg_' + name, SELECTION, missingStatsValue=-9999)
    cols = {}
    for (i, c) in enumerate(rat.cols):
        if c[0] != 'Histogram':
            cols[c[0]] = numpy.asarray(c[3])
    meta = {'selection': [list(s) for s in SELECTION], 'imgNull': imgNull, 'missing': -9999,
        'columns': list(cols.keys())}
    path = os.path.join(HERE, 'stats_%s.npz' % name)
    numpy.savez_compressed(path, seg=seg, img=img, meta=json.dumps(meta),
        **dict(('col_' + k, v) for (k, v) in cols.items()))
    print(name, seg.shape, int(seg.max()), 'segments', dict((k, (v.dtype.name, v[1:4].tolist())) for (k, v) in cols.items()))


if __name__ == '__main__':
    run('u16', 300, 340, numpy.uint16, 12, 1, 0.0, None)
    run('u16_null', 300, 340, numpy.uint16, 12, 2, 0.1, 0)
    run('u8', 200, 260, numpy.uint8, 9, 3, 0.0, 255)
    run('i16', 200, 220, numpy.int16, 10, 4, 0.05, -32768)
    run('u32', 150, 180, numpy.uint32, 8, 5, 0.0, None)
    run('i32', 150, 180, numpy.int32, 8, 6, 0.05, -2**31)
