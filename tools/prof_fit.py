"""Like prof_tile.py but with centres from a scikit-learn fit (what the benchmark uses)."""
import sys, os, ctypes
import numpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyshepseg_b200 import shepseg, synth, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bands = int(sys.argv[2]) if len(sys.argv) > 2 else 4
img = synth.synth_tiled(n, n, bands, seed=1)
km = shepseg.fitSpectralClusters(numpy.ascontiguousarray(img[:, ::6, ::6]), 60, 100, None, True)
for i in range(2):
    res = shepseg.doShepherdSegmentation(img, minSegmentSize=50, kmeansObj=km)
    tm = res.timings
    print('%dx%dx%d fitted centres: dev total %.2f assign %.2f clump %.2f single %.2f small %.2f | segs %d' % (
        n, n, bands, tm['total'], tm['assign'], tm['clump'], tm['single'], tm['small'], res.segimg.max()), flush=True)
ctx = _lib.default_context()
ctx.call('ssg_profile_enable', 1)
shepseg.doShepherdSegmentation(img, minSegmentSize=50, kmeansObj=km)
buf = ctypes.create_string_buffer(1 << 16)
ctx.call('ssg_profile_fetch', buf, len(buf))
rows_ = [l.split() for l in buf.value.decode().splitlines()]
tot = sum(float(r[2]) for r in rows_)
for r in sorted(rows_, key=lambda r: -float(r[2])):
    print('   %-34s x%-3s %8.3f ms  %5.1f%%' % (r[0], r[1], float(r[2]), 100 * float(r[2]) / tot))
print('   kernels total %.3f ms' % tot)
