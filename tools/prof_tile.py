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

    # per-kernel device time of one more run (events around every launch)
    import ctypes
    from pyshepseg_b200 import _lib
    ctx = _lib.default_context()
    ctx.call('ssg_profile_enable', 1)
    shepseg.doShepherdSegmentation(img, minSegmentSize=50, kmeansObj=km)
    buf = ctypes.create_string_buffer(1 << 16)
    ctx.call('ssg_profile_fetch', buf, len(buf))
    ctx.call('ssg_profile_enable', 0)
    rows_ = [l.split() for l in buf.value.decode().splitlines()]
    tot = sum(float(r[2]) for r in rows_)
    for r in sorted(rows_, key=lambda r: -float(r[2])):
        print('   %-34s x%-3s %8.3f ms  %5.1f%%' % (r[0], r[1], float(r[2]), 100 * float(r[2]) / tot))
    print('   kernels total %.3f ms' % tot)


if __name__ == '__main__':
    main()
