"""Time the per-segment statistics on a scene-sized label raster (device-resident and from host)."""
import sys
import time

import numpy

sys.path.insert(0, '.')
from pyshepseg_b200 import _lib, tilingstats  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10980
cell = int(sys.argv[2]) if len(sys.argv) > 2 else 24
rng = numpy.random.default_rng(0)
# blocky segments of about cell x cell pixels with ragged edges
(yy, xx) = numpy.mgrid[0:n, 0:n].astype(numpy.int32)
jit = rng.integers(0, cell // 3 + 1, (2, n, n), dtype=numpy.int32)
seg = (((yy + jit[0]) // cell) * ((n + cell) // cell + 1) + (xx + jit[1]) // cell + 1).astype(numpy.uint32)
del yy, xx, jit
(u, inv) = numpy.unique(seg, return_inverse=True)
seg = (inv.reshape(n, n) + 1).astype(numpy.uint32)
img = rng.integers(0, 4000, (n, n)).astype(numpy.uint16)
sel = [('mean', 'mean'), ('std', 'stddev'), ('med', 'median'), ('mode', 'mode'), ('p25', 'percentile', 25),
    ('n', 'pixcount'), ('min', 'min'), ('max', 'max')]
ctx = _lib.Context(0)
maxSegId = int(seg.max())
print('pixels %d segments %d' % (seg.size, maxSegId))
for rep in range(3):
    t0 = time.perf_counter()
    cols = tilingstats.calcPerSegmentStats(img, seg, sel, context=ctx)
    print('host arrays   : %.1f ms' % ((time.perf_counter() - t0) * 1e3))
dseg = ctx.dev_alloc(seg.nbytes)
dimg = ctx.dev_alloc(img.nbytes)
ctx.h2d(dseg, seg)
ctx.h2d(dimg, img)
ctx.lib.ssg_profile_enable(ctx.h, 1)
for rep in range(3):
    t0 = time.perf_counter()
    cols2 = tilingstats.calcPerSegmentStats(dimg, dseg, sel, maxSegId=maxSegId, context=ctx, shape=seg.shape, dtype=img.dtype)
    print('device arrays : %.1f ms' % ((time.perf_counter() - t0) * 1e3))
buf = (b' ' * 65536)
import ctypes
cbuf = ctypes.create_string_buffer(1 << 16)
ctx.lib.ssg_profile_fetch(ctx.h, cbuf, len(cbuf))
print(cbuf.value.decode())
for k in cols:
    assert numpy.array_equal(cols[k], cols2[k])
print('mean[1:4]', cols['mean'][1:4], 'std', cols['std'][1:4])
