"""One tile through ssg_segment_tile a few times (profiling target for ncu)."""
import sys, os
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyshepseg_b200 import shepseg, synth

def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
    bands = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    img = synth.synth_tiled(rows, cols, bands, seed=1)
    class KM: pass
    km = KM(); km.cluster_centers_ = synth.diagonal_centres(img, 60)
    for i in range(reps):
        res = shepseg.doShepherdSegmentation(img, minSegmentSize=50, kmeansObj=km)
        tm = res.timings
        print('%dx%dx%d: dev total %.2f assign %.2f clump %.2f single %.2f small %.2f | clumps %d segs %d' % (
            rows, cols, bands, tm['total'], tm['assign'], tm['clump'], tm['single'], tm['small'],
            tm['numClumps'], res.segimg.max()), flush=True)

if __name__ == '__main__':
    main()
