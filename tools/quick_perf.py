"""Stage timing probe (device milliseconds per stage) for a few tile shapes."""
import sys, time, os
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyshepseg_b200 import shepseg, synth, _lib

def run(rows, cols, bands, k, minSeg, reps=3):
    img = synth.synth_tiled(rows, cols, bands, seed=1)
    centres = synth.diagonal_centres(img, k)
    class KM: pass
    km = KM(); km.cluster_centers_ = centres
    for i in range(reps):
        t = time.time()
        res = shepseg.doShepherdSegmentation(img, minSegmentSize=minSeg, kmeansObj=km)
        wall = time.time() - t
        tm = res.timings
        print('%dx%dx%d k=%d minSeg=%d: wall %.1f ms | dev total %.2f assign %.2f clump %.2f single %.2f small %.2f | clumps %d segs %d passes %d rounds %d | %.1f Mpix/s dev' % (
            rows, cols, bands, k, minSeg, wall * 1e3, tm['total'], tm['assign'], tm['clump'], tm['single'], tm['small'],
            tm['numClumps'], res.segimg.max(), tm['numSmallPasses'], tm['numSinglePixelRounds'],
            rows * cols / tm['total'] / 1e3), flush=True)

if __name__ == '__main__':
    run(1000, 1000, 3, 60, 50)
    run(4096, 4096, 4, 60, 50)
    run(4096, 4096, 10, 60, 50)
    if len(sys.argv) > 1 and sys.argv[1] == 'big':
        run(7908, 7908, 4, 60, 50, reps=2)
        run(8000, 8000, 6, 30, 100, reps=2)
