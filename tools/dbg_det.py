"""Determinism probe: one tile several times, CRC of the labels and the per-size merge counts."""
import os, sys, zlib
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyshepseg_b200 import shepseg, synth
n = int(sys.argv[1]); reps = int(sys.argv[2])
img = synth.synth_tiled(n, n, 4, seed=1)
class KM: pass
km = KM(); km.cluster_centers_ = synth.diagonal_centres(img, 60)
for i in range(reps):
    r = shepseg.doShepherdSegmentation(img, minSegmentSize=50, kmeansObj=km)
    print('CRC %08x segs %d elim %d small-ms %.2f' % (zlib.crc32(r.segimg) & 0xffffffff, r.segimg.max(), r.smallSegmentsEliminated, r.timings['small']), flush=True)
