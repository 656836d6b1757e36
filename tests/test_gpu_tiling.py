"""
GPU parity tests of the tiled path: pyshepseg_b200.tiling.doTiledShepherdSegmentation against
the mosaics the unmodified reference wrote (tests/golden/tiled_*.npz) and against the CPU
oracle's per-tile segmentation + stitch on larger seeded rasters.  Bit-exact label rasters,
histograms and maxSegId.
"""
import numpy
import pytest

import goldenutil
from oracle import oracle
from pyshepseg_b200 import synth, rasterfile

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def tiling():
    from pyshepseg_b200 import tiling as mod
    return mod


def same(got, want, what):
    got = numpy.asarray(got)
    want = numpy.asarray(want)
    assert got.shape == want.shape, '%s: shape %s != %s' % (what, got.shape, want.shape)
    if not numpy.array_equal(got, want):
        bad = numpy.argwhere(got != want)
        first = tuple(bad[0])
        raise AssertionError('%s: %d of %d differ; first at %s: got %s want %s' % (
            what, len(bad), got.size, first, got[first], want[first]))


def run_tiled(tiling, img, km, m, cfg=None, nullVal='meta'):
    src = rasterfile.MemoryRaster(img, nodata=m['imgNullVal'])
    res = tiling.doTiledShepherdSegmentation(src, None, tileSize=m['tileSize'],
        overlapSize=m['overlapSize'], minSegmentSize=m['minSegmentSize'],
        numClusters=m['numClusters'], imgNullVal=m['imgNullVal'],
        fourConnected=m['fourConnected'], kmeansObj=km,
        simpleTileRecode=m['simpleTileRecode'], outputDriver='MEM', returnGDALDS=True,
        concurrencyCfg=cfg)
    return res


@pytest.mark.parametrize('name', goldenutil.tiled_names())
def test_golden_tiled(tiling, name):
    c = goldenutil.load(name)
    m = c['meta']
    res = run_tiled(tiling, c['img'], goldenutil.Centres(c['centres']), m)
    assert res.numTileRows == m['numTileRows'] and res.numTileCols == m['numTileCols']
    same(res.outDs.array, c['mosaic'], 'mosaic')
    assert int(res.maxSegId) == m['maxSegId']
    same(res.outDs.hist, c['hist'], 'histogram')
    assert float(res.maxSpectralDiff) == m['maxSpectralDiff']
    assert res.outDs.nodata == 0
    assert res.outDs.metadata['LAYER_TYPE'] == 'thematic'
    for name in ('walltime', 'reading', 'segmentation', 'stitchtiles'):
        assert res.timings.getDurationsForName(name) is not None


@pytest.mark.parametrize('name', ['tiled_700x900', 'tiled_640_5x5_8conn'])
def test_threads_equal_sequential(tiling, name):
    c = goldenutil.load(name)
    m = c['meta']
    cfg = tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS, numWorkers=3)
    res = run_tiled(tiling, c['img'], goldenutil.Centres(c['centres']), m, cfg)
    same(res.outDs.array, c['mosaic'], 'mosaic (3 workers)')
    assert int(res.maxSegId) == m['maxSegId']
    same(res.outDs.hist, c['hist'], 'histogram (3 workers)')
    assert res.timings.getDurationsForName('startworkers') is not None


def oracle_tiled(img, km, tileSize, overlap, minSeg, nullVal, four, simple=False):
    (nB, nR, nC) = img.shape
    ti = oracle.getTilesForFile(nC, nR, tileSize, overlap)
    segs = {}
    for ((col, row), (x, y, xs, ys)) in ti.tiles.items():
        sub = numpy.ascontiguousarray(img[:, y:y + ys, x:x + xs])
        segs[(col, row)] = oracle.doShepherdSegmentation(sub, minSegmentSize=minSeg, imgNullVal=nullVal,
            fourConnected=four, kmeansObj=km).segimg
    return oracle.stitchTiles(segs, ti, nC, nR, overlap, simpleTileRecode=simple)


CASES = [
    ('wide_2000x2600', (2000, 2600, 4), 512, 128, 40, 40, 0.0, True),
    ('null_1500x1700_8conn', (1500, 1700, 3), 400, 100, 30, 30, 0.15, False),
    ('small_overlap', (1300, 1200, 3), 300, 20, 20, 25, 0.0, True),
    # the tile layout of BASELINE config 2 (10980 with 4096/1024: 2 x 2 tiles, the last row and
    # column grown to 7908) at a quarter of the size: 2745 with 1024/256 -> tiles of 1024 and 1977
    ('c2_layout_quarter', (2745, 2745, 4), 1024, 256, 60, 50, 0.0, True),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_tiled_against_oracle(tiling, case):
    (name, (r, c, b), tileSize, overlap, k, minSeg, nullFrac, four) = case
    img = synth.synth_v1(r, c, b, seed=len(name) + 40, nullFrac=nullFrac)
    nullVal = 0 if nullFrac > 0 else None
    km = goldenutil.Centres(synth.diagonal_centres(img, k, nullVal))
    (mosaic, maxSegId, hist) = oracle_tiled(img, km, tileSize, overlap, minSeg, nullVal, four)
    m = {'tileSize': tileSize, 'overlapSize': overlap, 'minSegmentSize': minSeg, 'numClusters': k,
        'imgNullVal': nullVal, 'fourConnected': four, 'simpleTileRecode': False}
    res = run_tiled(tiling, img, km, m)
    same(res.outDs.array, mosaic, name + ' mosaic')
    assert int(res.maxSegId) == maxSegId
    same(res.outDs.hist, hist, name + ' histogram')
    # the tiled partition equals the whole-image partition away from nothing: every label id
    # 1..maxSegId that has pixels is connected-consistent with the oracle (checked by equality)


def test_file_to_file(tiling, tmp_path):
    """GeoTIFF in, GeoTIFF out, through the built-in reader / writer; null value from the file"""
    img = synth.synth_v1(900, 1000, 4, seed=77, nullFrac=0.1)
    km = goldenutil.Centres(synth.diagonal_centres(img, 20, 0))
    infile = str(tmp_path / 'in.tif')
    outfile = str(tmp_path / 'out.tif')
    rasterfile.writeImage(infile, img, nodata=0)
    res = tiling.doTiledShepherdSegmentation(infile, outfile, tileSize=400, overlapSize=100,
        minSegmentSize=30, kmeansObj=km, outputDriver='GTiff')
    (mosaic, maxSegId, hist) = oracle_tiled(img, km, 400, 100, 30, 0, True)
    out = rasterfile.openRaster(outfile)
    same(out.readWindow([1], 0, 0, 1000, 900)[0], mosaic, 'file mosaic')
    assert out.nodata == [0]
    assert int(res.maxSegId) == maxSegId
    same(numpy.load(outfile + '.hist.npy'), hist, 'file histogram')
    out.close()


def test_whole_file_kmeans_fit(tiling):
    """kmeansObj=None: the subsample fit runs on the host (scikit-learn) and the result carries
    it; determinism of the segmentation given those centres is checked against the oracle"""
    img = synth.synth_v1(800, 900, 3, seed=5)
    src = rasterfile.MemoryRaster(img)
    res = tiling.doTiledShepherdSegmentation(src, None, tileSize=512, overlapSize=64,
        minSegmentSize=30, numClusters=12, fixedKMeansInit=True, outputDriver='MEM',
        returnGDALDS=True)
    assert res.kmeans.cluster_centers_.shape == (12, 3)
    assert 0 < res.subsamplePcnt <= 100
    (mosaic, maxSegId, hist) = oracle_tiled(img, res.kmeans, 512, 64, 30, None, True)
    same(res.outDs.array, mosaic, 'mosaic with fitted centres')
    assert res.timings.getDurationsForName('spectralclusters') is not None


def test_worker_failure_surfaces_at_once(tiling):
    """A worker that dies (here: the raster read raises) must make the call fail quickly with the
    real cause, not with a tile-completion timeout (tiling.py:926-928, checkWorkerExceptions)"""
    import time

    class Broken(rasterfile.MemoryRaster):
        def readWindow(self, *a, **k):
            raise IOError('disk on fire')
    img = synth.synth_v1(600, 600, 3, seed=3)
    src = Broken(img)
    src.img = None        # not addressable in place: every tile goes through readWindow
    km = goldenutil.Centres(synth.diagonal_centres(img, 10))
    cfg = tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS, numWorkers=2,
        tileCompletionTimeout=30)
    t0 = time.time()
    with pytest.raises(tiling.PyShepSegTilingError, match='disk on fire'):
        tiling.doTiledShepherdSegmentation(src, None, tileSize=256, overlapSize=64, minSegmentSize=20,
            kmeansObj=km, outputDriver='MEM', concurrencyCfg=cfg)
    assert time.time() - t0 < 20


def test_unsupported_raster_type_is_a_tiling_error(tiling):
    img = numpy.zeros((2, 64, 64), dtype=numpy.float32)
    km = goldenutil.Centres(numpy.zeros((3, 2)))
    with pytest.raises(tiling.PyShepSegTilingError, match='not supported'):
        tiling.doTiledShepherdSegmentation(rasterfile.MemoryRaster(img), None, tileSize=32, overlapSize=8,
            kmeansObj=km, outputDriver='MEM')


@pytest.mark.parametrize('workers', [0, 2])
def test_overviews_cut_on_the_device(tiling, workers):
    """The overview levels of every written window as TilingSegmenter.writeOverviews makes them
    (tiling.py:1360-1383: arr[L//2::L, L//2::L] of the trimmed window at (xOff//L, yOff//L)),
    cut out by ssg_window_overviews; the default MEM output carries the reference's level list."""
    c = goldenutil.load('tiled_700x900')
    m = c['meta']
    img = c['img']
    (nB, nR, nC) = img.shape
    levels = [2, 4, 8, 5]
    sink = rasterfile.MemorySink(nC, nR, levels=levels)
    cfg = tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS, numWorkers=workers) \
        if workers else None
    res = tiling.doTiledShepherdSegmentation(rasterfile.MemoryRaster(img, nodata=m['imgNullVal']), sink,
        tileSize=m['tileSize'], overlapSize=m['overlapSize'], minSegmentSize=m['minSegmentSize'],
        numClusters=m['numClusters'], imgNullVal=m['imgNullVal'], fourConnected=m['fourConnected'],
        kmeansObj=goldenutil.Centres(c['centres']), simpleTileRecode=m['simpleTileRecode'], returnGDALDS=True,
        concurrencyCfg=cfg)
    same(sink.array, c['mosaic'], 'mosaic')
    ti = tiling.getTilesForFile((nC, nR), m['tileSize'], m['overlapSize'])
    for lvl in levels:
        want = numpy.zeros((-(-nR // lvl), -(-nC // lvl)), dtype=numpy.uint32)
        for ((col, row), (x, y, xs, ys)) in sorted(ti.tiles.items(), key=lambda kv: (kv[0][1], kv[0][0])):
            (top, bottom, left, right) = tiling.tileMargins(ti, col, row, xs, ys, m['overlapSize'])
            arr = c['mosaic'][y + top:y + bottom, x + left:x + right]
            sub = arr[lvl // 2::lvl, lvl // 2::lvl]
            (xo, yo) = ((x + left) // lvl, (y + top) // lvl)
            sub = sub[:want.shape[0] - yo, :want.shape[1] - xo]
            want[yo:yo + sub.shape[0], xo:xo + sub.shape[1]] = sub
        same(sink.overviews[lvl], want, 'overview level %d' % lvl)
    # the output the function creates itself has the reference's levels (none for a raster this small)
    res2 = run_tiled(tiling, img, goldenutil.Centres(c['centres']), m)
    assert list(res2.outDs.levels) == rasterfile.overviewLevels(nC, nR)

