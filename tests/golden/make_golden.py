"""
Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(pyshepseg at /root/reference, under the numba / scikit-learn / scipy / numpy installed
in the development image).  Run from the repo root:

    python tests/golden/make_golden.py

The fixtures pin the CPU oracle (tests/test_oracle_golden.py) and, on a GPU box, the
CUDA path (tests/test_gpu_parity.py).  /root/reference does not exist on the GPU box,
so no GPU test, smoke() or bench.py imports it; besides this script only
tests/test_oracle_vs_reference_live.py does (a CPU test that is skipped where the reference
checkout does not exist).

Every .npz holds the input image, the cluster centres given to both implementations,
the call parameters and the reference's outputs after each stage of
doShepherdSegmentation (shepseg.py:130-249): clusters, clump labels, labels after
eliminateSinglePixels, final labels, and the counters of the SegmentationResult.
"""
import json
import os
import sys

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, '/root/reference')

import fake_gdal  # noqa: E402
fake_gdal.install()

from pyshepseg import shepseg, tiling  # noqa: E402
from pyshepseg_b200 import synth  # noqa: E402


def versions():
    import numba
    import sklearn
    import scipy
    return {'numpy': numpy.__version__, 'numba': numba.__version__,
        'sklearn': sklearn.__version__, 'scipy': scipy.__version__,
        'pyshepseg': getattr(sys.modules['pyshepseg'], '__version__', '?')}


def fit(img, numClusters, imgNullVal):
    return shepseg.fitSpectralClusters(img, numClusters, 100, imgNullVal, True)


class FixedCentres(object):
    """Wraps a fitted KMeans, replacing its centres (used for integer-centre ties)."""
    def __init__(self, km, centres):
        import copy
        self.km = copy.deepcopy(km)
        self.km.cluster_centers_ = numpy.ascontiguousarray(centres, dtype=numpy.float64)
        self.cluster_centers_ = self.km.cluster_centers_

    def predict(self, x):
        return self.km.predict(x)


def run_stages(img, km, minSegmentSize, maxSpectralDiff, imgNullVal, fourConnected,
        spectDistPcntile=50):
    """The body of shepseg.doShepherdSegmentation, keeping every intermediate."""
    out = {}
    clusters = shepseg.applySpectralClusters(km, img, imgNullVal)
    out['clusters'] = clusters.astype(numpy.int32)
    (seg, maxSegId) = shepseg.clump(clusters, shepseg.SEGNULLVAL, fourConnected=fourConnected,
        clumpId=shepseg.MINSEGID)
    maxSegId = shepseg.SegIdType(maxSegId - 1)
    out['clumps'] = seg.copy()
    out['numClumps'] = int(maxSegId)
    segSize = shepseg.makeSegSize(seg)
    oldMaxSegId = maxSegId
    shepseg.eliminateSinglePixels(img, seg, segSize, shepseg.MINSEGID, maxSegId, fourConnected)
    out['seg_singles'] = seg.copy()
    maxSegId = seg.max()
    out['singlePixelsEliminated'] = int(oldMaxSegId - maxSegId)
    msd = shepseg.autoMaxSpectralDiff(km, maxSpectralDiff, spectDistPcntile)
    out['msd_value'] = numpy.float64(msd)
    out['msd_is_f32'] = int(isinstance(msd, numpy.float32))
    numElimSmall = shepseg.eliminateSmallSegments(seg, img, maxSegId, minSegmentSize, msd,
        fourConnected, shepseg.MINSEGID)
    out['seg_final'] = seg.copy()
    out['smallSegmentsEliminated'] = int(numElimSmall)
    # cross-check against the top-level entry point
    res = shepseg.doShepherdSegmentation(img, minSegmentSize=minSegmentSize,
        maxSpectralDiff=maxSpectralDiff, imgNullVal=imgNullVal, fourConnected=fourConnected,
        kmeansObj=km, spectDistPcntile=spectDistPcntile)
    assert numpy.array_equal(res.segimg, seg)
    assert int(res.singlePixelsEliminated) == out['singlePixelsEliminated']
    assert int(res.smallSegmentsEliminated) == out['smallSegmentsEliminated']
    return out


def save_case(name, img, km, params, out):
    meta = dict(params)
    meta['versions'] = versions()
    msd = params['maxSpectralDiff']
    meta['maxSpectralDiff'] = msd if (msd is None or isinstance(msd, (str, int))) else float(msd)
    path = os.path.join(HERE, name + '.npz')
    numpy.savez_compressed(path, img=img, centres=numpy.asarray(km.cluster_centers_, numpy.float64),
        meta=json.dumps(meta), **out)
    print('%-28s %s  clumps=%d singles=%d small=%d final=%d  (%d kB)' % (name, img.shape,
        out['numClumps'], out['singlePixelsEliminated'], out['smallSegmentsEliminated'],
        int(out['seg_final'].max()), os.path.getsize(path) // 1024))


def single_tile_cases():
    cases = []

    def add(name, img, k, minSeg, msd, nullVal, four, pcntile=50, centres=None):
        cases.append((name, img, k, minSeg, msd, nullVal, four, pcntile, centres))

    base = synth.synth_v1(150, 190, 3, seed=11, cell=16)
    add('v1_4conn_auto', base, 12, 20, 'auto', None, True)
    add('v1_8conn_auto', base, 12, 20, 'auto', None, False)
    add('v1_4conn_msd300', base, 12, 20, 300.0, None, True)
    add('v1_4conn_msdnone', base, 12, 20, None, None, True)
    add('v1_4conn_msdint', base, 12, 20, 250, None, True)
    add('v1_minseg1', base, 12, 1, 'auto', None, True)
    add('v1_pcntile20', base, 12, 30, 'auto', None, False, 20)
    nullimg = synth.synth_v1(140, 170, 4, seed=12, cell=16, nullFrac=0.1, nullVal=0)
    add('v1_null_4conn', nullimg, 10, 25, 'auto', 0, True)
    add('v1_null_8conn', nullimg, 10, 25, 'auto', 0, False)
    add('v1_10band', synth.synth_v1(120, 130, 10, seed=13, cell=16), 16, 20, 'auto', None, True)
    add('v1_6band_k30', synth.synth_v1(160, 160, 6, seed=14, cell=24, nullFrac=0.1),
        30, 40, 'auto', 0, True)
    q = (synth.synth_v1(170, 180, 3, seed=15, cell=16) // 64 * 64).astype(numpy.uint16)
    add('v1_quantised', q, 12, 30, 'auto', None, True)
    add('v1_quantised_8conn', q, 12, 30, None, None, False)
    big = synth.synth_v1(120, 140, 3, seed=16, cell=16, lo=30000, hi=62000, noise=400)
    add('v1_bigvalues', big, 12, 60, None, None, True)
    u8 = synth.synth_v1(130, 150, 3, seed=17, cell=16, lo=20, hi=230, noise=6, dtype=numpy.uint8)
    add('v1_uint8', u8, 10, 20, 'auto', None, True)
    i16 = (synth.synth_v1(130, 150, 3, seed=18, cell=16, lo=200, hi=3000).astype(numpy.int32) -
        1500).astype(numpy.int16)
    add('v1_int16_negative', i16, 10, 20, 'auto', None, True)
    one = synth.synth_v1(90, 100, 3, seed=19, cell=16)
    one[:, 40, 50] = 0
    add('one_null_pixel', one, 8, 15, 'auto', 0, True)
    two = one.copy()
    two[:, 10, 12] = 0
    add('two_null_pixels', two, 8, 15, 'auto', 0, True)
    partial = synth.synth_v1(90, 100, 3, seed=20, cell=16)
    partial[1, 20:40, 30:60] = 0      # null in one band only
    add('null_in_one_band', partial, 8, 15, 'auto', 0, False)
    flat = synth.synth_flat(260, 280, 3, numCells=5, seed=21, border=4, nullVal=65535)
    add('flat_cap_4conn', flat, 5, 50, 'auto', 65535, True)
    add('flat_cap_8conn', flat, 5, 50, 'auto', 65535, False)
    noisyflat = flat.copy()
    rng = numpy.random.default_rng(22)
    m = rng.random(flat.shape[1:]) < 0.02
    noisyflat[:, m] = rng.integers(100, 60000, (3, int(m.sum())))
    add('flat_cap_speckled', noisyflat, 8, 30, None, 65535, True)
    tie = synth.synth_v1(100, 110, 2, seed=23, cell=16, lo=0, hi=40, noise=3, dtype=numpy.uint8)
    tieCentres = numpy.array([[10, 10], [20, 10], [10, 20], [20, 20], [30, 30], [15, 15], [15, 15]],
        dtype=numpy.float64)
    add('integer_centre_ties', tie, 7, 10, 'auto', None, True, 50, tieCentres)
    return cases


def make_single_tile():
    for (name, img, k, minSeg, msd, nullVal, four, pcntile, centres) in single_tile_cases():
        km = fit(img, k, nullVal)
        if centres is not None:
            km = FixedCentres(km, centres)
        out = run_stages(img, km, minSeg, msd, nullVal, four, pcntile)
        params = {'numClusters': k, 'minSegmentSize': minSeg, 'maxSpectralDiff': msd,
            'imgNullVal': nullVal, 'fourConnected': four, 'spectDistPcntile': pcntile}
        save_case(name, img, km, params, out)


def make_clump_only():
    """clump() on its own: the cap boundary strips of SURVEY probe A5 and flat blocks."""
    out = {}
    for n in (9999, 10000, 10001, 10002, 10005, 20003):
        img = numpy.ones((1, n), dtype=numpy.int32)
        for four in (True, False):
            (lab, nxt) = shepseg.clump(img, 0, four, 1)
            out['strip_%d_%d' % (n, int(four))] = numpy.array(
                [nxt] + list(numpy.bincount(lab.ravel())[1:]), dtype=numpy.int64)
    for (shape, four) in (((200, 200), True), ((200, 200), False), ((97, 311), True),
            ((97, 311), False), ((311, 97), False)):
        img = numpy.full(shape, 3, dtype=numpy.int32)
        (lab, nxt) = shepseg.clump(img, 0, four, 1)
        out['flat_%dx%d_%d' % (shape[0], shape[1], int(four))] = lab
    rng = numpy.random.default_rng(5)
    img = rng.integers(0, 3, (180, 220)).astype(numpy.int32)
    img[40:150, 30:200] = 2
    img[60:70, 50:60] = 0
    for four in (True, False):
        (lab, nxt) = shepseg.clump(img, 0, four, 7)
        out['mixed_%d' % int(four)] = lab
        out['mixed_%d_next' % int(four)] = numpy.int64(nxt)
    out['mixed_img'] = img
    path = os.path.join(HERE, 'clump_only.npz')
    numpy.savez_compressed(path, **out)
    print('clump_only.npz  (%d kB)' % (os.path.getsize(path) // 1024))


def make_tiled():
    """doTiledShepherdSegmentation through the fake GDAL: the mosaic and its tiles."""
    cases = [
        ('tiled_700x900', synth.synth_v1(700, 900, 3, seed=31, cell=16), 256, 64, 12, 30, None, True, False),
        ('tiled_null_4x4', synth.synth_v1(620, 660, 3, seed=32, cell=24, nullFrac=0.1), 180, 60, 10, 40, 0, True, False),
        ('tiled_640_5x5_8conn', synth.synth_v1(640, 640, 3, seed=33, cell=16), 200, 120, 10, 25, None, False, False),
        ('tiled_simple_recode', synth.synth_v1(500, 620, 3, seed=34, cell=16), 256, 64, 10, 25, None, True, True),
    ]
    for (name, img, tileSize, overlap, k, minSeg, nullVal, four, simple) in cases:
        km = fit(img, k, nullVal)
        fake_gdal.put_image('in_' + name, img, nodata=nullVal)
        res = tiling.doTiledShepherdSegmentation('in_' + name, 'out_' + name, tileSize=tileSize,
            overlapSize=overlap, minSegmentSize=minSeg, numClusters=k, imgNullVal=nullVal,
            fourConnected=four, kmeansObj=km, simpleTileRecode=simple)
        outDs = fake_gdal.REGISTRY['out_' + name]
        band = outDs.GetRasterBand(1)
        mosaic = band.arr.copy()
        hist = band.rat.cols[0][3]
        meta = {'tileSize': tileSize, 'overlapSize': overlap, 'numClusters': k,
            'minSegmentSize': minSeg, 'imgNullVal': nullVal, 'fourConnected': four,
            'simpleTileRecode': simple, 'maxSegId': int(res.maxSegId),
            'numTileRows': int(res.numTileRows), 'numTileCols': int(res.numTileCols),
            'maxSpectralDiff': float(res.maxSpectralDiff), 'versions': versions(),
            'overviewLevels': [int(band.overviews[i].arr.shape[0]) for i in range(len(band.overviews))]}
        path = os.path.join(HERE, name + '.npz')
        numpy.savez_compressed(path, img=img, centres=numpy.asarray(km.cluster_centers_, numpy.float64),
            mosaic=mosaic, hist=numpy.asarray(hist), meta=json.dumps(meta))
        print('%-28s %s tiles=%dx%d maxSegId=%d (%d kB)' % (name, img.shape, res.numTileRows,
            res.numTileCols, res.maxSegId, os.path.getsize(path) // 1024))


if __name__ == '__main__':
    what = sys.argv[1:] or ['single', 'clump', 'tiled']
    if 'single' in what:
        make_single_tile()
    if 'clump' in what:
        make_clump_only()
    if 'tiled' in what:
        make_tiled()
