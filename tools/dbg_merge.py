"""Debug helper: one seeded case through the GPU path and the oracle, several times; prints how many
pixels differ (between GPU runs and against the oracle)."""
import os, sys
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyshepseg_b200 import shepseg, synth
from oracle import oracle

def main():
    (r, c, b, k, minSeg, nullFrac) = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]),
        int(sys.argv[5]), float(sys.argv[6]))
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
    img = synth.synth_tiled(r, c, b, seed=26)
    nullVal = None
    if nullFrac > 0:
        rr = numpy.arange(r)[:, None]; cc = numpy.arange(c)[None, :]
        img[:, (rr + cc) < numpy.sqrt(2.0 * nullFrac * r * c)] = 0
        nullVal = 0
    class KM: pass
    km = KM(); km.cluster_centers_ = synth.diagonal_centres(img, k, nullVal)
    want = oracle.doShepherdSegmentation(img, numClusters=k, minSegmentSize=minSeg, imgNullVal=nullVal, kmeansObj=km)
    prev = None
    for i in range(reps):
        got = shepseg.doShepherdSegmentation(img, numClusters=k, minSegmentSize=minSeg, imgNullVal=nullVal, kmeansObj=km)
        d = int((got.segimg != want.segimg).sum())
        d2 = -1 if prev is None else int((got.segimg != prev).sum())
        print('%dx%dx%d k=%d minSeg=%d null=%.2f run %d: vs oracle %d differ, vs previous run %d; segs %d / %d; elim %d / %d' % (
            r, c, b, k, minSeg, nullFrac, i, d, d2, got.segimg.max(), want.segimg.max(),
            got.smallSegmentsEliminated, want.smallSegmentsEliminated), flush=True)
        prev = got.segimg

main()
