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
