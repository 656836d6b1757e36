"""Host-side raster I/O used by the tiled driver (no GPU needed)."""
import os

import numpy
import pytest

from pyshepseg_b200 import rasterfile, tiling, timinghooks
from oracle import oracle


@pytest.mark.parametrize('dtype', [numpy.uint8, numpy.uint16, numpy.int16])
def test_tiff_roundtrip_windows(tmp_path, dtype):
    rng = numpy.random.default_rng(3)
    info = numpy.iinfo(dtype)
    img = rng.integers(info.min, info.max, (3, 57, 83)).astype(dtype)
    fn = str(tmp_path / 'in.tif')
    rasterfile.writeImage(fn, img, nodata=7)
    src = rasterfile.TiffSource(fn)
    assert (src.count, src.ysize, src.xsize) == img.shape
    assert src.dtype == numpy.dtype(dtype)
    assert src.nodata == [7, 7, 7]
    win = src.readWindow([1, 3], 10, 20, 30, 15)
    assert numpy.array_equal(win, img[[0, 2], 20:35, 10:40])
    full = src.readWindow([1, 2, 3], 0, 0, 83, 57)
    assert numpy.array_equal(full, img)
    src.close()


def test_tiff_sink_is_readable(tmp_path):
    fn = str(tmp_path / 'out.tif')
    sink = rasterfile.createRaster(fn, 40, 30, 'GTiff', [])
    a = numpy.arange(12, dtype=numpy.uint32).reshape(3, 4) + 5
    sink.write(a, 7, 9)
    sink.setNoData(0)
    sink.writeHistogram(numpy.arange(4, dtype=numpy.float64))
    sink.close()
    src = rasterfile.TiffSource(fn)
    got = src.readWindow([1], 0, 0, 40, 30)[0]
    want = numpy.zeros((30, 40), dtype=numpy.uint32)
    want[9:12, 7:11] = a
    assert numpy.array_equal(got, want)
    assert src.nodata == [0]
    assert os.path.exists(fn + '.hist.npy')
    src.close()


def test_npy_source_and_sink(tmp_path):
    img = numpy.arange(2 * 5 * 6, dtype=numpy.uint16).reshape(2, 5, 6)
    fn = str(tmp_path / 'a.npy')
    numpy.save(fn, img)
    src = rasterfile.openRaster(fn)
    assert numpy.array_equal(src.readWindow([2], 1, 2, 3, 2)[0], img[1, 2:4, 1:4])
    out = str(tmp_path / 'b.npy')
    sink = rasterfile.createRaster(out, 6, 5, 'NPY', [])
    sink.write(numpy.ones((2, 2), dtype=numpy.uint32), 4, 3)
    sink.close()
    assert numpy.load(out)[3:, 4:].sum() == 4


def test_tile_layout_matches_reference_layouts():
    """SURVEY probe A13: the layouts getTilesForFile produces for the benchmark rasters"""
    for (size, tile, ov, n, last) in ((10980, 4096, 1024, 2, (3072, 7908)), (8000, 4096, 1024, 1, (0, 8000)),
            (40000, 4096, 1024, 12, (33792, 6208)), (1000, 4096, 1024, 1, (0, 1000)),
            (700, 256, 64, 2, (192, 508))):
        ti = tiling.getTilesForFile((size, size), tile, ov)
        assert (ti.ncols, ti.nrows) == (n, n)
        assert ti.getTile(n - 1, n - 1)[0::2] == last
        ref = oracle.getTilesForFile(size, size, tile, ov)
        assert ref.tiles == ti.tiles


def test_mode_by_key():
    keys = numpy.array([5, 5, 5, 9, 9, 2, 2, 2])
    vals = numpy.array([7, 3, 7, 1, 4, 6, 8, 6])
    cnts = numpy.array([2, 4, 2, 1, 1, 1, 5, 1])
    (k, m) = tiling._modeByKey(keys, vals, cnts)
    got = dict(zip(k.tolist(), m.tolist()))
    # 5: value 7 has 4, value 3 has 4 -> smallest (3); 9: tie 1/1 -> 1; 2: 8 has 5
    assert got == {5: 3, 9: 1, 2: 8}


def test_timers():
    t = timinghooks.Timers()
    with t.interval('a'):
        pass
    with t.interval('a'):
        pass
    d = t.makeSummaryDict()
    assert d['a']['count'] == 2 and d['a']['total'] >= 0
    import pickle
    t2 = pickle.loads(pickle.dumps(t))
    assert len(t2.getDurationsForName('a')) == 2


def test_argument_errors():
    img = numpy.zeros((3, 10, 10), dtype=numpy.uint16)
    with pytest.raises(tiling.PyShepSegTilingError):
        tiling.doTiledShepherdSegmentation(img, None, overlapSize=3, outputDriver='MEM')
    with pytest.raises(tiling.PyShepSegTilingError):
        tiling.doTiledShepherdSegmentation(img, 'x.kea', outputDriver='NoSuchDriver')
    with pytest.raises(ValueError):
        tiling.doTiledShepherdSegmentation(img, None, outputDriver='MEM',
            concurrencyCfg=tiling.SegmentationConcurrencyConfig(concurrencyType='bogus'))


def test_overview_levels_follow_the_reference_loop():
    """setupOverviews (tiling.py:1385-1404) appends a level, re-tests the same level and only then
    moves on: one level more than a plain `while size // level >= 1024` gives."""
    assert rasterfile.overviewLevels(10980, 10980) == [4, 8, 16]
    assert rasterfile.overviewLevels(4096, 100) == [4, 8]
    assert rasterfile.overviewLevels(4000, 3000) == []
    assert rasterfile.overviewLevels(40000, 40000) == [4, 8, 16, 32, 64]


def test_memory_sink_overviews_and_caller_array():
    """writeOverviews takes arr[L//2::L, L//2::L] at (xoff//L, yoff//L) (tiling.py:1360-1383)"""
    full = numpy.arange(40 * 64, dtype=numpy.uint32).reshape(40, 64)
    mine = numpy.zeros((40, 64), dtype=numpy.uint32)
    sink = rasterfile.MemorySink(64, 40, array=mine, levels=[4, 8])
    for (y, x) in ((0, 0), (0, 32), (24, 0), (24, 32)):
        win = full[y:y + 24, x:x + 32]
        sink.write(win, x, y)
        sink.writeOverviews(win, x, y)
    assert sink.array is mine and numpy.array_equal(mine, full)
    assert numpy.array_equal(sink.overviews[4], full[2::4, 2::4])
    assert numpy.array_equal(sink.overviews[8], full[4::8, 4::8])
    with pytest.raises(rasterfile.RasterError):
        rasterfile.MemorySink(10, 10, array=mine)


def test_memory_raster_row_band_view():
    img = numpy.arange(2 * 6 * 5, dtype=numpy.uint16).reshape(2, 6, 5)
    band = rasterfile.MemoryRaster(img[:, 2:5], yoff=2, fullYsize=6)
    assert (band.ysize, band.fullYsize, band.yoff) == (3, 6, 2)
    assert rasterfile.MemoryRaster(img).fullYsize == 6


def test_statistics_from_histogram_equal_the_reference_expressions():
    """utils.estimateStatsFromHisto (utils.py:47-95), restated with fewer passes: same strings"""
    def reference(hist):
        mask = hist > 0
        nVals = hist.sum()
        minVal = mask.argmax()
        maxVal = hist.shape[0] - numpy.flip(mask).argmax() - 1
        values = numpy.arange(hist.shape[0])
        meanVal = (values * hist).sum() / nVals
        stdDevVal = numpy.sqrt((hist * numpy.power(values - meanVal, 2)).sum() / nVals)
        medianVal = (hist.cumsum() >= hist.sum() / 2).nonzero()[0][0]
        return [repr(int(minVal)), repr(int(maxVal)), repr(float(meanVal)), repr(float(stdDevVal)),
            repr(int(numpy.argmax(hist))), repr(int(medianVal))]
    rng = numpy.random.default_rng(1)
    for n in (5, 1000, 200001):
        for trial in range(3):
            hist = rng.integers(0, 3000, n).astype(numpy.float64)
            hist[0] = 0
            if trial == 1:
                hist[:n // 3] = 0
            if trial == 2:
                hist[-(n // 4):] = 0
                hist[1] = 7
            sink = rasterfile.MemorySink(4, 4)
            tiling.estimateStatsFromHisto(sink, hist)
            m = sink.metadata
            assert [m['STATISTICS_' + k] for k in ('MINIMUM', 'MAXIMUM', 'MEAN', 'STDDEV', 'MODE', 'MEDIAN')] == \
                reference(hist)
