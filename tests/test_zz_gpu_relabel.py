"""
shepseg.relabelSegments on the GPU against the oracle's restatement of shepseg.py:739-777
(order-preserving squeeze of the ids that own no pixel; in place; segSize left untouched).
"""
import numpy
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('minSegId', [1, 3])
def test_relabel_segments(minSegId):
    from pyshepseg_b200 import shepseg
    rng = numpy.random.default_rng(5 + minSegId)
    seg = rng.integers(0, 4000, (300, 417)).astype(numpy.uint32)
    seg[rng.random(seg.shape) < 0.3] = 0
    # remove a third of the ids so that there are gaps to squeeze out
    gone = rng.random(4000) < 0.33
    gone[:minSegId] = False
    seg[gone[seg]] = 7 if not gone[7] else 0
    segSize = numpy.bincount(seg.ravel(), minlength=4000).astype(numpy.uint32)
    want = seg.copy()
    oracle.relabelSegments(want, segSize, minSegId)
    got = seg.copy()
    sizeBefore = segSize.copy()
    shepseg.relabelSegments(got, segSize, minSegId)
    assert numpy.array_equal(got, want)
    assert numpy.array_equal(segSize, sizeBefore)
    # contiguous from minSegId up
    present = numpy.unique(got)
    present = present[present >= minSegId]
    assert numpy.array_equal(present, numpy.arange(minSegId, minSegId + len(present)))
