"""BASELINE.json configs 1, 3 and 5 at full size on one GPU (timing; parity is in tests/)."""
import sys, os, time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyshepseg_b200 import shepseg, tiling, synth, rasterfile, _lib


class KM(object):
    pass


def c1():
    img = synth.synth_v1(1000, 1000, 3, seed=0)
    km = KM(); km.cluster_centers_ = synth.diagonal_centres(img, 60)
    for i in range(3):
        t = time.time()
        res = shepseg.doShepherdSegmentation(img, numClusters=60, minSegmentSize=50, kmeansObj=km)
        wall = time.time() - t
    print('C1 1000x1000x3 k=60 minSeg=50: wall %.1f ms (%.0f Mpix/s), device %.2f ms, %d segments' % (
        wall * 1e3, 1.0 / wall, res.timings['total'], res.segimg.max()), flush=True)


def c3():
    (r, c, b) = (8000, 8000, 6)
    pin = _lib.PinnedArray((b, r, c), numpy.uint16)
    img = synth.synth_tiled(r, c, b, seed=2, out=pin.array)
    rr = numpy.arange(r)[:, None]; cc = numpy.arange(c)[None, :]
    img[:, (rr + cc) < numpy.sqrt(2.0 * 0.10 * r * c)] = 0
    km = KM(); km.cluster_centers_ = synth.diagonal_centres(img, 30, 0)
    for i in range(3):
        t = time.time()
        res = shepseg.doShepherdSegmentation(img, numClusters=30, minSegmentSize=100, imgNullVal=0, kmeansObj=km)
        wall = time.time() - t
    tm = res.timings
    print('C3 8000x8000x6 null wedge k=30 minSeg=100: wall %.1f ms (%.0f Mpix/s host to host), device %.2f ms '
        '(assign %.2f clump %.2f single %.2f small %.2f), %d segments' % (wall * 1e3, r * c / wall / 1e6, tm['total'],
        tm['assign'], tm['clump'], tm['single'], tm['small'], res.segimg.max()), flush=True)
    pin.free()


def c5():
    (r, c, b) = (10980, 10980, 10)
    pin = _lib.PinnedArray((b, r, c), numpy.uint16)
    img = synth.synth_tiled(r, c, b, seed=4, out=pin.array)
    km = KM(); km.cluster_centers_ = synth.diagonal_centres(img, 60)
    cfg = tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS, numWorkers=3,
        tileCompletionTimeout=600)
    for i in range(3):
        t = time.time()
        res = tiling.doTiledShepherdSegmentation(rasterfile.MemoryRaster(img), None, tileSize=4096, overlapSize=1024,
            minSegmentSize=50, maxSpectralDiff='auto', spectDistPcntile=50, kmeansObj=km, outputDriver='MEM',
            returnGDALDS=True, concurrencyCfg=cfg)
        wall = time.time() - t
    print('C5 10980x10980x10 tiled 4096/1024 auto: wall %.1f ms (%.0f Mpix/s host to host), %d segments, stages %s' % (
        wall * 1e3, r * c / wall / 1e6, res.maxSegId, dict((k, round(v, 1)) for (k, v) in res.stageMs.items())), flush=True)
    pin.free()


if __name__ == '__main__':
    c1(); c3(); c5()
