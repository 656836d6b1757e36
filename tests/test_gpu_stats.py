"""
GPU parity tests of the per-segment statistics (pyshepseg_b200.tilingstats, csrc/stats.cu):
every column bit for bit against the RAT columns the unmodified reference wrote
(tests/golden/stats_*.npz, made by tests/golden/make_golden_stats.py), against the numpy oracle
on fresh seeded cases, and through the file level entry point with a stand-in GDAL.
"""
import glob
import json
import os
import sys

import numpy
import pytest

from oracle import stats_oracle

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, 'golden', 'stats_*.npz')))

SELECTION = [('b_min', 'min'), ('b_max', 'max'), ('b_mean', 'mean'), ('b_std', 'stddev'),
    ('b_median', 'median'), ('b_mode', 'mode'), ('b_p25', 'percentile', 25), ('b_p75', 'percentile', 75),
    ('b_p0', 'percentile', 0), ('b_p100', 'percentile', 100), ('b_count', 'pixcount')]


@pytest.fixture(scope='module')
def tilingstats():
    from pyshepseg_b200 import tilingstats as mod
    return mod


def load(path):
    z = numpy.load(path)
    meta = json.loads(str(z['meta']))
    cols = dict((k[4:], z[k]) for k in z.files if k.startswith('col_'))
    return (z['seg'], z['img'], meta, cols)


def same(got, want):
    assert set(got) == set(want)
    for (name, col) in want.items():
        assert got[name].dtype == col.dtype, name
        if not numpy.array_equal(got[name], col):
            bad = numpy.flatnonzero(got[name] != col)
            raise AssertionError('%s: %d segments differ, first %d: got %r want %r' % (
                name, len(bad), bad[0], got[name][bad[0]], col[bad[0]]))


def segments(nR, nC, cell, seed, nullFrac=0.0):
    """Irregular segments: nearest seed point labels, numbered 1.. without gaps."""
    rng = numpy.random.default_rng(seed)
    n = max(2, (nR // cell) * (nC // cell))
    py = rng.integers(0, nR, n)
    px = rng.integers(0, nC, n)
    (yy, xx) = numpy.mgrid[0:nR, 0:nC]
    best = numpy.full((nR, nC), numpy.inf)
    lab = numpy.zeros((nR, nC), dtype=numpy.int64)
    for i in range(n):
        d = (yy - py[i]) ** 2 + (xx - px[i]) ** 2
        m = d < best
        best[m] = d[m]
        lab[m] = i + 1
    if nullFrac:
        lab[rng.random((nR, nC)) < nullFrac] = 0
    (u, inv) = numpy.unique(lab, return_inverse=True)
    lut = numpy.arange(len(u)) + (1 if u[0] != 0 else 0)
    return lut[inv].reshape(nR, nC).astype(numpy.uint32)


@pytest.mark.parametrize('path', GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_stats_equal_reference_fixture(tilingstats, path):
    (seg, img, meta, want) = load(path)
    sel = [tuple(s) for s in meta['selection']]
    got = tilingstats.calcPerSegmentStats(img, seg, sel, meta['missing'], meta['imgNull'])
    same(got, want)
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize('dtype,imgNull,seed', [
    (numpy.uint8, None, 11), (numpy.uint8, 0, 12), (numpy.uint16, 65535, 13), (numpy.int16, -1, 14),
    (numpy.uint32, None, 15), (numpy.int32, 7, 16)])
def test_stats_equal_oracle(tilingstats, dtype, imgNull, seed):
    rng = numpy.random.default_rng(seed)
    (nR, nC) = (310, 421)
    seg = segments(nR, nC, 14, seed, nullFrac=0.03)
    info = numpy.iinfo(dtype)
    # few distinct values in some segments (ties for the mode), the full range in others
    wide = rng.integers(info.min, int(info.max) + 1, (nR, nC), dtype=numpy.int64)
    narrow = rng.integers(0, 5, (nR, nC), dtype=numpy.int64)
    img = numpy.where(seg % 3 == 0, narrow, wide).astype(dtype)
    if imgNull is not None:
        img[rng.random((nR, nC)) < 0.1] = imgNull
        img[seg == 5] = imgNull
    want = stats_oracle.calcPerSegmentStats(img, seg, SELECTION, -9999, imgNull)
    got = tilingstats.calcPerSegmentStats(img, seg, SELECTION, -9999, imgNull)
    same(got, want)


def test_stats_large_segments_and_subset_of_statistics(tilingstats):
    """Few large segments (long histograms per thread), a selection without float columns."""
    rng = numpy.random.default_rng(3)
    seg = segments(700, 900, 230, 3)
    img = rng.integers(0, 4000, seg.shape).astype(numpy.uint16)
    sel = [('m', 'median'), ('n', 'pixcount'), ('q', 'percentile', 99)]
    same(tilingstats.calcPerSegmentStats(img, seg, sel), stats_oracle.calcPerSegmentStats(img, seg, sel, -9999, None))
    sel = [('s', 'stddev')]
    same(tilingstats.calcPerSegmentStats(img, seg, sel), stats_oracle.calcPerSegmentStats(img, seg, sel, -9999, None))


def test_stats_on_resident_labels(tilingstats):
    """Device pointers: labels and image that are already in GPU memory."""
    from pyshepseg_b200 import _lib
    (seg, img, meta, want) = load(GOLDEN[0])
    ctx = _lib.Context(0)
    try:
        dseg = ctx.dev_alloc(seg.nbytes)
        dimg = ctx.dev_alloc(img.nbytes)
        ctx.h2d(dseg, seg)
        ctx.h2d(dimg, img)
        sel = [tuple(s) for s in meta['selection']]
        got = tilingstats.calcPerSegmentStats(dimg, dseg, sel, meta['missing'], meta['imgNull'],
            maxSegId=int(seg.max()), context=ctx, shape=seg.shape, dtype=img.dtype)
        same(got, want)
        ctx.dev_free(dseg)
        ctx.dev_free(dimg)
    finally:
        ctx.close()


def test_stats_segment_id_out_of_range(tilingstats):
    from pyshepseg_b200 import _lib
    seg = numpy.array([[1, 2], [3, 9]], dtype=numpy.uint32)
    img = numpy.ones((2, 2), dtype=numpy.uint8)
    with pytest.raises(_lib.ShepsegB200Error, match='above maxSegId'):
        tilingstats.calcPerSegmentStats(img, seg, [('a', 'min')], maxSegId=3)


# ---- the file level entry point -----------------------------------------------------------
@pytest.fixture()
def fakegdal(monkeypatch):
    sys.path.insert(0, os.path.join(HERE, 'golden'))
    import fake_gdal
    saved = dict((k, sys.modules.get(k)) for k in ('osgeo', 'osgeo.gdal', 'osgeo.gdal_array', 'osgeo.osr'))
    gdal = fake_gdal.install()
    yield (fake_gdal, gdal)
    for (k, v) in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    sys.path.remove(os.path.join(HERE, 'golden'))


def put(fake_gdal, gdal, name, seg, img, imgNull, hist=True):
    fake_gdal.put_image('img_' + name, img[None], nodata=imgNull)
    segds = fake_gdal.put_image('seg_' + name, seg[None].copy())
    rat = segds.GetRasterBand(1).GetDefaultRAT()
    if hist:
        segSize = numpy.bincount(seg.ravel(), minlength=int(seg.max()) + 1).astype(numpy.float64)
        segSize[0] = 0
        rat.SetRowCount(len(segSize))
        rat.CreateColumn('Histogram', gdal.GFT_Real, gdal.GFU_PixelCount)
        rat.WriteArray(segSize, 0)
    return (segds, rat)


@pytest.mark.parametrize('path', GOLDEN[:2], ids=[os.path.basename(p)[:-4] for p in GOLDEN[:2]])
def test_tiled_stats_write_the_rat_through_gdal(tilingstats, fakegdal, path):
    (fake_gdal, gdal) = fakegdal
    (seg, img, meta, want) = load(path)
    sel = [tuple(s) for s in meta['selection']]
    (segds, rat) = put(fake_gdal, gdal, 't', seg, img, meta['imgNull'])
    res = tilingstats.calcPerSegmentStatsTiled('img_t', 1, 'seg_t', sel, missingStatsValue=meta['missing'])
    assert res.timings is not None
    got = dict((c[0], numpy.asarray(c[3])) for c in rat.cols if c[0] != 'Histogram')
    same(got, want)
    types = dict((c[0], c[1]) for c in rat.cols)
    assert types['b_mean'] == gdal.GFT_Real and types['b_std'] == gdal.GFT_Real
    assert types['b_min'] == gdal.GFT_Integer and types['b_count'] == gdal.GFT_Integer
    # a second run reuses the columns
    ncols = len(rat.cols)
    tilingstats.calcPerSegmentStatsTiled('img_t', 1, segds, sel, missingStatsValue=meta['missing'])
    assert len(rat.cols) == ncols


def test_tiled_stats_errors(tilingstats, fakegdal):
    (fake_gdal, gdal) = fakegdal
    (seg, img, meta, want) = load(GOLDEN[0])
    sel = [('a', 'mean')]
    put(fake_gdal, gdal, 'e', seg, img, None, hist=False)
    with pytest.raises(tilingstats.PyShepSegStatsError, match='Histogram column must exist'):
        tilingstats.calcPerSegmentStatsTiled('img_e', 1, 'seg_e', sel)
    put(fake_gdal, gdal, 'f', seg, img.astype(numpy.float32), None)
    with pytest.raises(tilingstats.PyShepSegStatsError, match='Float image types'):
        tilingstats.calcPerSegmentStatsTiled('img_f', 1, 'seg_f', sel)
    put(fake_gdal, gdal, 'g', seg, img, None)
    fake_gdal.put_image('img_g', img[None, :-1])
    with pytest.raises(tilingstats.PyShepSegStatsError, match='same size'):
        tilingstats.calcPerSegmentStatsTiled('img_g', 1, 'seg_g', sel)
    (segds, rat) = put(fake_gdal, gdal, 'h', seg, img, None)
    fake_gdal.REGISTRY['img_h'].SetGeoTransform((5.0, 1.0, 0.0, 0.0, 0.0, -1.0))
    with pytest.raises(tilingstats.PyShepSegStatsError, match='same spatial extent'):
        tilingstats.calcPerSegmentStatsTiled('img_h', 1, 'seg_h', sel)
    # a Histogram column that disagrees with the raster: segments never complete
    (segds, rat) = put(fake_gdal, gdal, 'i', seg, img, None)
    rat.cols[0][3][3] += 1
    with pytest.raises(tilingstats.PyShepSegStatsError, match='Not all pixels found'):
        tilingstats.calcPerSegmentStatsTiled('img_i', 1, 'seg_i', sel)
    with pytest.raises(KeyError):
        tilingstats.calcPerSegmentStatsTiled('img_i', 1, 'seg_i', [('a', 'variance')])


def test_tiled_stats_builtin_formats(tilingstats, tmp_path):
    """Without GDAL: label raster + Histogram side file as the built-in sinks write them."""
    from pyshepseg_b200 import rasterfile
    (seg, img, meta, want) = load(GOLDEN[1])
    segfile = str(tmp_path / 'seg.npy')
    imgfile = str(tmp_path / 'img.npy')
    sink = rasterfile.createRaster(segfile, seg.shape[1], seg.shape[0], 'NPY', None)
    sink.write(seg, 0, 0)
    hist = numpy.bincount(seg.ravel(), minlength=int(seg.max()) + 1).astype(numpy.float64)
    hist[0] = 0
    sink.writeHistogram(hist)
    sink.close()
    numpy.save(imgfile, img[None])
    json.dump({'nodata': meta['imgNull']}, open(imgfile + '.json', 'w'))
    sel = [tuple(s) for s in meta['selection']]
    res = tilingstats.calcPerSegmentStatsTiled(imgfile, 1, segfile, sel, missingStatsValue=meta['missing'])
    same(res.columns, want)
    same(tilingstats.readRat(segfile), want)
