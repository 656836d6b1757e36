"""The clump cap at the scale of the reference's own test image (cmdline/runtests.py:145-265: flat Voronoi
cells far over MAX_CLUMP_SIZE): GPU result against the oracle, stage times."""
import os, sys, time
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyshepseg_b200 import shepseg, synth
from oracle import oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 100
img = synth.synth_flat(n, n, 3, numCells=cells, seed=3, border=0, nullVal=65535)
class KM: pass
km = KM(); km.cluster_centers_ = synth.diagonal_centres(img, cells, 65535)
t0 = time.time()
want = oracle.doShepherdSegmentation(img, numClusters=cells, minSegmentSize=50, imgNullVal=65535, kmeansObj=km)
tOracle = time.time() - t0
for i in range(2):
    t0 = time.time()
    got = shepseg.doShepherdSegmentation(img, numClusters=cells, minSegmentSize=50, imgNullVal=65535, kmeansObj=km)
    wall = time.time() - t0
    tm = got.timings
    print('%dx%dx3 flat, %d cells: GPU wall %.1f ms, dev total %.1f (assign %.1f clump %.1f single %.1f small %.1f), oversized regions %d, clumps %d, segments %d | oracle %.1f s | equal %s' % (
        n, n, cells, wall * 1e3, tm['total'], tm['assign'], tm['clump'], tm['single'], tm['small'], tm['numOversized'],
        tm['numClumps'], got.segimg.max(), tOracle, bool(numpy.array_equal(got.segimg, want.segimg))), flush=True)
