"""
BASELINE.json configs 3 and 5 at FULL size on the GPU against the CPU oracle (bit-exact):
  C3  8000 x 8000 x 6 uint16, 10 % null wedge, numClusters=30, minSegmentSize=100, one tile
      (doShepherdSegmentation);
  C5  10980 x 10980 x 10 uint16, tiled 4096 / 1024, maxSpectralDiff='auto', spectDistPcntile=50,
      minSegmentSize=50 (doTiledShepherdSegmentation).
Config 2 (10980 x 10980 x 4, the benchmark) is checked at full size by bench.py itself after its
timed regions (`parity_vs_oracle`), and at the layout's quarter size in test_gpu_tiling.py.
The oracle runs every tile on its own host thread (the C port releases the GIL).
"""
import threading

import numpy
import pytest

import goldenutil
from oracle import oracle
from pyshepseg_b200 import synth, rasterfile

pytestmark = pytest.mark.gpu


def same(got, want, what):
    if not numpy.array_equal(got, want):
        bad = numpy.argwhere(numpy.asarray(got) != numpy.asarray(want))
        first = tuple(bad[0])
        raise AssertionError('%s: %d of %d differ; first at %s: got %s want %s' % (
            what, len(bad), numpy.asarray(got).size, first, got[first], want[first]))


def test_c3_landsat_like_8000_nulls():
    from pyshepseg_b200 import shepseg
    (nR, nC, nB) = (8000, 8000, 6)
    img = synth.synth_tiled(nR, nC, nB, seed=2)
    # upper-left null wedge of 10 % of the pixels
    lim = numpy.sqrt(2 * 0.10 * nR * nC)
    rr = numpy.arange(nR)[:, None]
    cc = numpy.arange(nC)[None, :]
    img[:, (rr + cc) < lim] = 0
    km = goldenutil.Centres(synth.diagonal_centres(img, 30, 0))
    want = [None]

    def cpu():
        want[0] = oracle.doShepherdSegmentation(img, numClusters=30, minSegmentSize=100, imgNullVal=0, kmeansObj=km)
    th = threading.Thread(target=cpu)
    th.start()
    got = shepseg.doShepherdSegmentation(img, numClusters=30, minSegmentSize=100, imgNullVal=0, kmeansObj=km)
    th.join()
    same(got.segimg, want[0].segimg, 'C3 labels')
    assert int(got.singlePixelsEliminated) == int(want[0].singlePixelsEliminated)
    assert got.smallSegmentsEliminated == want[0].smallSegmentsEliminated
    assert float(got.maxSpectralDiff) == float(want[0].maxSpectralDiff)
    assert (got.segimg[:100, :100] == 0).all()


def test_c5_sentinel2_stack_10band_tiled():
    from pyshepseg_b200 import tiling
    (nR, nC, nB) = (10980, 10980, 10)
    img = synth.synth_tiled(nR, nC, nB, seed=4)
    km = goldenutil.Centres(synth.diagonal_centres(img, 60))
    ti = oracle.getTilesForFile(nC, nR, 4096, 1024)
    assert (ti.nrows, ti.ncols) == (2, 2)
    segs = {}

    def cpu(cr):
        (x, y, xs, ys) = ti.tiles[cr]
        sub = numpy.ascontiguousarray(img[:, y:y + ys, x:x + xs])
        segs[cr] = oracle.doShepherdSegmentation(sub, minSegmentSize=50, maxSpectralDiff='auto',
            spectDistPcntile=50, kmeansObj=km).segimg
    ths = [threading.Thread(target=cpu, args=(cr,)) for cr in
        sorted(ti.tiles, key=lambda cr: -ti.tiles[cr][2] * ti.tiles[cr][3])]
    for t in ths:
        t.start()
    cfg = tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS, numWorkers=2)
    res = tiling.doTiledShepherdSegmentation(rasterfile.MemoryRaster(img), None, tileSize=4096,
        overlapSize=1024, minSegmentSize=50, numClusters=60, maxSpectralDiff='auto', spectDistPcntile=50,
        kmeansObj=km, outputDriver='MEM', returnGDALDS=True, concurrencyCfg=cfg)
    for t in ths:
        t.join()
    (mosaic, maxSegId, hist) = oracle.stitchTiles(segs, ti, nC, nR, 1024)
    same(res.outDs.array, mosaic, 'C5 mosaic')
    assert int(res.maxSegId) == maxSegId
    same(res.outDs.hist, hist, 'C5 histogram')
