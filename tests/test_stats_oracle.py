"""The numpy restatement of the per-segment statistics against the fixtures written by the
unmodified reference (tests/golden/make_golden_stats.py): every column, bit for bit."""
import glob
import json
import os

import numpy
import pytest

from oracle import stats_oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stats_*.npz')))


def load(path):
    z = numpy.load(path)
    meta = json.loads(str(z['meta']))
    cols = dict((k[4:], z[k]) for k in z.files if k.startswith('col_'))
    return (z['seg'], z['img'], meta, cols)


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_stats_oracle_equals_reference(path):
    (seg, img, meta, want) = load(path)
    sel = [tuple(s) for s in meta['selection']]
    got = stats_oracle.calcPerSegmentStats(img, seg, sel, meta['missing'], meta['imgNull'])
    assert len(GOLDEN) >= 4
    for (name, col) in want.items():
        assert got[name].dtype == col.dtype, name
        assert numpy.array_equal(got[name], col), '%s differs at %s' % (name, numpy.flatnonzero(got[name] != col)[:5])


# ---- host logic of pyshepseg_b200.tilingstats (no GPU needed: everything here fails or returns
# before the library is touched) -----------------------------------------------------------------
def test_stats_selection_and_argument_errors():
    from pyshepseg_b200 import tilingstats
    # the reference's STATID_* numbers (tilingstats.py:770-777)
    assert tilingstats.statIDdict == {'min': 0, 'max': 1, 'mean': 2, 'stddev': 3, 'median': 4, 'mode': 5,
        'percentile': 6, 'pixcount': 7}
    (ids, params) = tilingstats.checkStatsSelection([('a', 'mean'), ('b', 'percentile', 25), ('c', 'pixcount')])
    assert ids.tolist() == [2, 6, 7] and params.tolist() == [0, 25, 0]
    with pytest.raises(KeyError):
        tilingstats.checkStatsSelection([('a', 'variance')])
    with pytest.raises(IndexError):
        tilingstats.checkStatsSelection([('a', 'percentile')])
    seg = numpy.ones((4, 5), dtype=numpy.uint32)
    with pytest.raises(tilingstats.PyShepSegStatsError, match='Float image types'):
        tilingstats.calcPerSegmentStats(numpy.ones((4, 5), dtype=numpy.float32), seg, [('a', 'mean')])
    with pytest.raises(tilingstats.PyShepSegStatsError, match='same size'):
        tilingstats.calcPerSegmentStats(numpy.ones((4, 6), dtype=numpy.uint8), seg, [('a', 'mean')])
    with pytest.raises(tilingstats.PyShepSegStatsError, match='not supported'):
        tilingstats.calcPerSegmentStats(numpy.ones((4, 5), dtype=numpy.int64), seg, [('a', 'mean')])
    with pytest.raises(ValueError, match='device pointers'):
        tilingstats.calcPerSegmentStats(12345, 67890, [('a', 'mean')])
    assert tilingstats.checkHistColumn(['x', 'Histogram']) == 1
    with pytest.raises(tilingstats.PyShepSegStatsError, match='Histogram column must exist'):
        tilingstats.checkHistColumn(['x', 'y'])
    assert tilingstats.equalProjection('abc', 'abc')


def test_stats_oracle_missing_and_percentile_edges():
    """an all-nodata segment: missing value everywhere but pixcount (0); percentile 0 is the LARGEST
    value (the reference's loop never runs and reads pixVals[-1]); percentile 100 the largest too"""
    seg = numpy.array([[1, 1, 1, 2, 2, 2]], dtype=numpy.uint32)
    img = numpy.array([[7, 3, 5, 0, 0, 0]], dtype=numpy.uint8)
    sel = [('lo', 'percentile', 0), ('hi', 'percentile', 100), ('med', 'median'), ('n', 'pixcount'), ('m', 'mean')]
    got = stats_oracle.calcPerSegmentStats(img, seg, sel, -1, 0)
    assert got['lo'].tolist() == [0, 7, -1] and got['hi'].tolist() == [0, 7, -1]
    assert got['med'].tolist() == [0, 5, -1] and got['n'].tolist() == [0, 3, 0]
    assert got['m'].tolist() == [0.0, 5.0, -1.0]
