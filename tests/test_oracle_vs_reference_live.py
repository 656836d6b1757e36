"""
Live differential test of the CPU oracle against the UNMODIFIED reference, on seeded random
cases that are not among the committed fixtures.  Runs only where the reference checkout exists
(/root/reference in the development container: it is absent on the GPU box, where the committed
fixtures of tests/golden/ do the pinning); the reference runs in a subprocess so that its numba
compilation and sys.path do not touch the test process.
"""
import json
import os
import subprocess
import sys

import numpy
import pytest

from oracle import oracle
from pyshepseg_b200 import synth
import goldenutil

REFERENCE = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    # rows, cols, bands, seed, k, minSeg, four, nullFrac, msd
    (120, 150, 3, 101, 12, 10, True, 0.0, 'auto'),
    (140, 110, 4, 102, 14, 15, False, 0.08, 'auto'),
    (100, 160, 3, 103, 10, 8, True, 0.0, None),
    (130, 130, 5, 104, 16, 12, False, 0.0, 250.0),
]

RUNNER = r'''
import json, sys
import numpy
sys.path.insert(0, %(ref)r)
from pyshepseg import shepseg
spec = json.load(open(sys.argv[1]))
for (i, c) in enumerate(spec['cases']):
    d = numpy.load(c['npz'])
    # the reference's own fit (scikit-learn); the centres it finds are handed to the oracle
    km = shepseg.fitSpectralClusters(d['img'], c['k'], 100, c['nullVal'], True)
    res = shepseg.doShepherdSegmentation(d['img'], numClusters=c['k'], minSegmentSize=c['minSeg'],
        maxSpectralDiff=c['msd'], imgNullVal=c['nullVal'], fourConnected=c['four'], kmeansObj=km)
    numpy.savez(c['out'], segimg=res.segimg, single=int(res.singlePixelsEliminated),
        small=int(res.smallSegmentsEliminated), msd=float(res.maxSpectralDiff),
        centres=numpy.asarray(km.cluster_centers_, dtype=numpy.float64))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'pyshepseg')),
    reason='the reference checkout is not on this machine')
def test_oracle_equals_reference_on_fresh_cases(tmp_path):
    try:
        import numba  # noqa: F401
        import sklearn  # noqa: F401
    except ImportError:
        pytest.skip('numba / scikit-learn missing: the reference cannot run here')
    cases = []
    inputs = []
    for (i, (r, c, b, seed, k, minSeg, four, nullFrac, msd)) in enumerate(CASES):
        img = synth.synth_v1(r, c, b, seed=seed, cell=12, nullFrac=nullFrac, nullVal=0)
        nullVal = 0 if nullFrac > 0 else None
        npz = str(tmp_path / ('in%d.npz' % i))
        numpy.savez(npz, img=img)
        cases.append({'npz': npz, 'out': str(tmp_path / ('out%d.npz' % i)), 'k': k, 'minSeg': minSeg,
            'msd': msd, 'nullVal': nullVal, 'four': four})
        inputs.append((img, minSeg, msd, nullVal, four))
    spec = str(tmp_path / 'spec.json')
    json.dump({'cases': cases}, open(spec, 'w'))
    script = str(tmp_path / 'run_reference.py')
    open(script, 'w').write(RUNNER % {'ref': REFERENCE})
    subprocess.run([sys.executable, script, spec], check=True, timeout=600, cwd=str(tmp_path))
    for (c, (img, minSeg, msd, nullVal, four)) in zip(cases, inputs):
        ref = numpy.load(c['out'])
        got = oracle.doShepherdSegmentation(img, minSegmentSize=minSeg, maxSpectralDiff=msd,
            imgNullVal=nullVal, fourConnected=four, kmeansObj=goldenutil.Centres(ref['centres']))
        assert numpy.array_equal(got.segimg, ref['segimg'])
        assert int(got.singlePixelsEliminated) == int(ref['single'])
        assert int(got.smallSegmentsEliminated) == int(ref['small'])
        assert float(got.maxSpectralDiff) == float(ref['msd'])


TILED_RUNNER = r'''
import json, sys
import numpy
sys.path.insert(0, %(golden)r)
sys.path.insert(0, %(ref)r)
import fake_gdal
fake_gdal.install()
from pyshepseg import shepseg, tiling
c = json.load(open(sys.argv[1]))
img = numpy.load(c['npz'])['img']
km = shepseg.fitSpectralClusters(img, c['k'], 100, c['nullVal'], True)
fake_gdal.put_image('live_in', img, nodata=c['nullVal'])
res = tiling.doTiledShepherdSegmentation('live_in', 'live_out', tileSize=c['tileSize'], overlapSize=c['overlap'],
    minSegmentSize=c['minSeg'], numClusters=c['k'], imgNullVal=c['nullVal'], fourConnected=c['four'], kmeansObj=km)
band = fake_gdal.REGISTRY['live_out'].GetRasterBand(1)
numpy.savez(c['out'], mosaic=band.arr, maxSegId=int(res.maxSegId),
    centres=numpy.asarray(km.cluster_centers_, dtype=numpy.float64))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'pyshepseg')),
    reason='the reference checkout is not on this machine')
def test_oracle_stitch_equals_reference_tiled_run(tmp_path):
    """tiling.doTiledShepherdSegmentation of the unmodified reference (through the numpy stand-in
    for GDAL) against the oracle's per-tile segmentation + stitchTiles, on a fresh raster"""
    try:
        import numba  # noqa: F401
        import sklearn  # noqa: F401
        import scipy  # noqa: F401
    except ImportError:
        pytest.skip('numba / scikit-learn / scipy missing: the reference cannot run here')
    img = synth.synth_v1(430, 520, 3, seed=211, cell=14, nullFrac=0.06, nullVal=0)
    c = {'npz': str(tmp_path / 'in.npz'), 'out': str(tmp_path / 'out.npz'), 'k': 11, 'nullVal': 0,
        'tileSize': 160, 'overlap': 48, 'minSeg': 14, 'four': False}
    numpy.savez(c['npz'], img=img)
    spec = str(tmp_path / 'spec.json')
    json.dump(c, open(spec, 'w'))
    script = str(tmp_path / 'run_reference_tiled.py')
    open(script, 'w').write(TILED_RUNNER % {'ref': REFERENCE, 'golden': os.path.join(ROOT, 'tests', 'golden')})
    subprocess.run([sys.executable, script, spec], check=True, timeout=900, cwd=str(tmp_path))
    ref = numpy.load(c['out'])
    km = goldenutil.Centres(ref['centres'])
    (nB, nR, nC) = img.shape
    ti = oracle.getTilesForFile(nC, nR, c['tileSize'], c['overlap'])
    segs = {}
    for ((col, row), (x, y, xs, ys)) in ti.tiles.items():
        sub = numpy.ascontiguousarray(img[:, y:y + ys, x:x + xs])
        segs[(col, row)] = oracle.doShepherdSegmentation(sub, minSegmentSize=c['minSeg'], imgNullVal=0,
            fourConnected=False, kmeansObj=km).segimg
    (mosaic, maxSegId, hist) = oracle.stitchTiles(segs, ti, nC, nR, c['overlap'])
    assert maxSegId == int(ref['maxSegId'])
    assert numpy.array_equal(mosaic, ref['mosaic'])


STATS_RUNNER = r'''
import json, os, sys
import numpy
sys.path.insert(0, %(golden)r)
import fake_gdal
gdal = fake_gdal.install()
class _SR(object):
    def __init__(self, wkt=''):
        self.wkt = wkt
    def IsSame(self, other):
        return self.wkt == other.wkt
sys.modules['osgeo.osr'].SpatialReference = _SR
sys.modules['osgeo.osr'].UseExceptions = lambda: None
sys.path.insert(0, %(ref)r)
from pyshepseg import tilingstats
spec = json.load(open(sys.argv[1]))
d = numpy.load(spec['npz'])
(seg, img) = (d['seg'], d['img'])
fake_gdal.put_image('img', img[None], nodata=spec['imgNull'])
segds = fake_gdal.put_image('seg', seg[None].copy())
rat = segds.GetRasterBand(1).GetDefaultRAT()
hist = numpy.bincount(seg.ravel(), minlength=int(seg.max()) + 1).astype(numpy.float64)
hist[0] = 0
rat.SetRowCount(len(hist))
rat.CreateColumn('Histogram', gdal.GFT_Real, gdal.GFU_PixelCount)
rat.WriteArray(hist, 0)
sel = [tuple(s) for s in spec['selection']]
tilingstats.calcPerSegmentStatsTiled('img', 1, 'seg', sel, missingStatsValue=spec['missing'])
numpy.savez(spec['out'], **dict((c[0], numpy.asarray(c[3])) for c in rat.cols if c[0] != 'Histogram'))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'pyshepseg')),
    reason='the reference checkout is not on this machine')
@pytest.mark.parametrize('dtype,imgNull,seed', [(numpy.uint16, None, 201), (numpy.int16, -5, 202), (numpy.uint32, 0, 203)])
def test_stats_oracle_equals_reference_on_fresh_cases(tmp_path, dtype, imgNull, seed):
    """the per-segment statistics: the unmodified tilingstats.calcPerSegmentStatsTiled (through the
    GDAL stand-in, in a subprocess) against oracle/stats_oracle.py on rasters that are not fixtures"""
    try:
        import numba  # noqa: F401
    except ImportError:
        pytest.skip('numba missing: the reference cannot run here')
    from oracle import stats_oracle
    rng = numpy.random.default_rng(seed)
    (nR, nC) = (150, 170)
    coarse = rng.integers(1, 400, (nR // 7 + 2, nC // 9 + 2))
    seg = numpy.kron(coarse, numpy.ones((7, 9), dtype=numpy.int64))[:nR, :nC]
    seg[rng.random((nR, nC)) < 0.04] = 0
    (_, inv) = numpy.unique(numpy.concatenate([[0], seg.ravel()]), return_inverse=True)
    seg = inv[1:].reshape(nR, nC).astype(numpy.uint32)       # ids 1..n without gaps, 0 = null
    info = numpy.iinfo(dtype)
    img = rng.integers(max(info.min, -30000), min(int(info.max), 3000000000) + 1, (nR, nC)).astype(dtype)
    img[seg % 4 == 0] = (img[seg % 4 == 0] % 7).astype(dtype)          # few distinct values: mode ties
    if imgNull is not None:
        img[rng.random((nR, nC)) < 0.1] = imgNull
        img[seg == 3] = imgNull
    sel = [('mn', 'min'), ('mx', 'max'), ('mean', 'mean'), ('sd', 'stddev'), ('med', 'median'), ('mode', 'mode'),
        ('p10', 'percentile', 10), ('p0', 'percentile', 0), ('p100', 'percentile', 100), ('n', 'pixcount')]
    npz = str(tmp_path / 'in.npz')
    out = str(tmp_path / 'out.npz')
    numpy.savez(npz, seg=seg, img=img)
    spec = {'npz': npz, 'out': out, 'imgNull': imgNull, 'missing': -9999, 'selection': [list(s) for s in sel]}
    specFile = str(tmp_path / 'spec.json')
    json.dump(spec, open(specFile, 'w'))
    runner = str(tmp_path / 'runner.py')
    open(runner, 'w').write(STATS_RUNNER % {'ref': REFERENCE, 'golden': os.path.join(ROOT, 'tests', 'golden')})
    env = dict(os.environ, NUMBA_CACHE_DIR=str(tmp_path / 'nbcache'))
    subprocess.run([sys.executable, runner, specFile], check=True, env=env, timeout=1200,
        stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    want = numpy.load(out)
    got = stats_oracle.calcPerSegmentStats(img, seg, sel, -9999, imgNull)
    for s in sel:
        assert got[s[0]].dtype == want[s[0]].dtype, s[0]
        assert numpy.array_equal(got[s[0]], want[s[0]]), '%s differs at %s' % (s[0],
            numpy.flatnonzero(got[s[0]] != want[s[0]])[:5])
