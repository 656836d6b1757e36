"""
Pins the CPU oracle (oracle/shepseg_oracle.c) against the golden fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  Every stage is compared bit for bit,
feeding each oracle stage the REFERENCE's output of the previous stage, and then the
whole chain end to end.
"""
import numpy
import pytest

from oracle import oracle
import goldenutil


@pytest.mark.parametrize('name', goldenutil.single_tile_names())
def test_stages_match_reference(name):
    c = goldenutil.load(name)
    m = c['meta']
    img = c['img']
    km = goldenutil.Centres(c['centres'])
    four = m['fourConnected']

    clusters = oracle.applySpectralClusters(km, img, m['imgNullVal'])
    assert clusters.dtype == numpy.int32
    assert numpy.array_equal(clusters, c['clusters'])

    (seg, nextId) = oracle.clump(c['clusters'], oracle.SEGNULLVAL, fourConnected=four,
        clumpId=oracle.MINSEGID)
    assert seg.dtype == numpy.uint32
    assert nextId - 1 == int(c['numClumps'])
    assert numpy.array_equal(seg, c['clumps'])

    seg = c['clumps'].copy()
    segSize = oracle.makeSegSize(seg)
    oracle.eliminateSinglePixels(img, seg, segSize, oracle.MINSEGID, nextId - 1, four)
    assert numpy.array_equal(seg, c['seg_singles'])
    assert int(c['numClumps']) - int(seg.max()) == int(c['singlePixelsEliminated'])

    msd = oracle.autoMaxSpectralDiff(km, goldenutil.msd_of(c), m['spectDistPcntile'])
    assert float(msd) == float(c['msd_value'])
    assert isinstance(msd, numpy.float32) == bool(c['msd_is_f32'])

    seg = c['seg_singles'].copy()
    numElim = oracle.eliminateSmallSegments(seg, img, seg.max(), m['minSegmentSize'], msd,
        four, oracle.MINSEGID)
    assert numElim == int(c['smallSegmentsEliminated'])
    assert numpy.array_equal(seg, c['seg_final'])


@pytest.mark.parametrize('name', goldenutil.single_tile_names())
def test_end_to_end_matches_reference(name):
    c = goldenutil.load(name)
    m = c['meta']
    res = oracle.doShepherdSegmentation(c['img'], minSegmentSize=m['minSegmentSize'],
        maxSpectralDiff=goldenutil.msd_of(c), imgNullVal=m['imgNullVal'],
        fourConnected=m['fourConnected'], kmeansObj=goldenutil.Centres(c['centres']),
        spectDistPcntile=m['spectDistPcntile'])
    assert res.segimg.dtype == numpy.uint32 and res.segimg.flags.c_contiguous
    assert numpy.array_equal(res.segimg, c['seg_final'])
    assert int(res.singlePixelsEliminated) == int(c['singlePixelsEliminated'])
    assert res.smallSegmentsEliminated == int(c['smallSegmentsEliminated'])
    assert res.numClumps == int(c['numClumps'])


def test_clump_cap_boundaries():
    """shepseg.py:481,502 -- the MAX_CLUMP_SIZE behaviour (SURVEY probe A5)."""
    g = numpy.load(goldenutil.GOLDEN_DIR + '/clump_only.npz')
    for key in g.files:
        if key.startswith('strip_'):
            (_, n, four) = key.split('_')
            img = numpy.ones((1, int(n)), dtype=numpy.int32)
            (lab, nxt) = oracle.clump(img, 0, bool(int(four)), 1)
            got = numpy.array([nxt] + list(numpy.bincount(lab.ravel())[1:]))
            assert numpy.array_equal(got, g[key]), key
        elif key.startswith('flat_'):
            (_, shape, four) = key.split('_')
            (r, cc) = shape.split('x')
            img = numpy.full((int(r), int(cc)), 3, dtype=numpy.int32)
            (lab, nxt) = oracle.clump(img, 0, bool(int(four)), 1)
            assert numpy.array_equal(lab, g[key]), key
    for four in (0, 1):
        (lab, nxt) = oracle.clump(g['mixed_img'], 0, bool(four), 7)
        assert numpy.array_equal(lab, g['mixed_%d' % four])
        assert nxt == int(g['mixed_%d_next' % four])


@pytest.mark.parametrize('name', goldenutil.tiled_names())
def test_tiled_matches_reference(name):
    """tiling.py:950-1306 -- per-tile oracle segmentation + oracle stitch reproduce the
    mosaic the reference's doTiledShepherdSegmentation wrote."""
    c = goldenutil.load(name)
    m = c['meta']
    img = c['img']
    km = goldenutil.Centres(c['centres'])
    (nBands, nRows, nCols) = img.shape
    tileInfo = oracle.getTilesForFile(nCols, nRows, m['tileSize'], m['overlapSize'])
    assert tileInfo.nrows == m['numTileRows'] and tileInfo.ncols == m['numTileCols']
    tileSegs = {}
    for ((col, row), (xpos, ypos, xsize, ysize)) in tileInfo.tiles.items():
        sub = numpy.ascontiguousarray(img[:, ypos:ypos + ysize, xpos:xpos + xsize])
        res = oracle.doShepherdSegmentation(sub, minSegmentSize=m['minSegmentSize'],
            imgNullVal=m['imgNullVal'], fourConnected=m['fourConnected'], kmeansObj=km)
        tileSegs[(col, row)] = res.segimg
    (mosaic, maxSegId, hist) = oracle.stitchTiles(tileSegs, tileInfo, nCols, nRows,
        m['overlapSize'], simpleTileRecode=m['simpleTileRecode'])
    assert maxSegId == m['maxSegId']
    assert numpy.array_equal(mosaic, c['mosaic'])
    refHist = c['hist']
    assert len(hist) == len(refHist)
    assert numpy.array_equal(hist, refHist)
