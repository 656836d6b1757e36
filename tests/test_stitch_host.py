"""
The stitch in its three-phase form (per-tile tables -> sequential resolve -> lut apply),
checked on the CPU: the per-tile tables that ssg_tile_tables_device computes on the GPU are
emulated here with numpy, the host part (pyshepseg_b200.tiling.resolveTile) is the product
code, and the result must equal the golden mosaics the reference wrote and the oracle's
restatement of tiling.stitchTiles.
"""
import numpy
import pytest

import goldenutil
from oracle import oracle
from pyshepseg_b200 import tiling, _lib


def numpy_tile_tables(tile, overlap, topB, leftB, top, bottom, left, right):
    """What ssg_tile_tables_device + ssg_tile_tables_fetch return, in numpy."""
    (ys, xs) = tile.shape
    maxId = int(tile.max())
    n = maxId + 1
    rr = numpy.repeat(numpy.arange(ys), xs)
    cc = numpy.tile(numpy.arange(xs), ys)
    flat = tile.ravel()
    big = numpy.iinfo(numpy.int64).max
    minRow = numpy.full(n, big)
    minCol = numpy.full(n, big)
    numpy.minimum.at(minRow, flat, rr)
    numpy.minimum.at(minCol, flat, cc)
    flags = numpy.zeros(n, dtype=numpy.uint8)
    present = minRow != big
    present[0] = False
    flags[present] |= _lib.SEG_PRESENT

    def crossing(strip, axisVals):
        mn = numpy.full(n, big)
        mx = numpy.full(n, -1)
        numpy.minimum.at(mn, strip.ravel(), axisVals.ravel())
        numpy.maximum.at(mx, strip.ravel(), axisVals.ravel())
        return (mn, mx)
    keyT = numpy.zeros(n, dtype=bool)
    keyL = numpy.zeros(n, dtype=bool)
    if topB is not None:
        strip = tile[:overlap, :]
        rows = numpy.repeat(numpy.arange(strip.shape[0]), xs).reshape(strip.shape)
        (mn, mx) = crossing(strip, rows)
        mid = strip.shape[0] // 2
        keyT = present & (mn < mid) & (mx >= mid)
    if leftB is not None:
        strip = tile[:, :overlap]
        cols = numpy.tile(numpy.arange(strip.shape[1]), ys).reshape(strip.shape)
        (mn, mx) = crossing(strip, cols)
        mid = strip.shape[1] // 2
        keyL = present & (mn < mid) & (mx >= mid)
    flags[keyT] |= _lib.SEG_KEYTOP
    flags[keyL] |= _lib.SEG_KEYLEFT
    inTrim = numpy.zeros(n, dtype=bool)
    inTrim[numpy.unique(tile[top:bottom, left:right])] = True
    inTrim[0] = False
    flags[inTrim] |= _lib.SEG_INTRIM
    numbered = (present & ~keyT & ~keyL & (minCol >= left) & (minRow >= top) &
        (minCol < right) & (minRow < bottom))
    flags[numbered] |= _lib.SEG_NUMBERED
    rank = numpy.zeros(n, dtype=numpy.uint32)
    rank[numbered] = numpy.arange(1, int(numbered.sum()) + 1)
    keys = []
    if topB is not None:
        a = tile[:overlap, :].ravel().astype(numpy.uint64)
        sel = keyT[tile[:overlap, :].ravel()]
        keys.append((a[sel] << numpy.uint64(32)) | topB.ravel()[sel].astype(numpy.uint64))
    if leftB is not None:
        a = tile[:, :overlap].ravel().astype(numpy.uint64)
        sel = keyL[tile[:, :overlap].ravel()]
        keys.append(numpy.uint64(_lib.PAIR_LEFT) | (a[sel] << numpy.uint64(32)) |
            leftB.ravel()[sel].astype(numpy.uint64))
    if keys:
        (pairKeys, pairCounts) = numpy.unique(numpy.concatenate(keys), return_counts=True)
    else:
        (pairKeys, pairCounts) = (numpy.zeros(0, numpy.uint64), numpy.zeros(0, numpy.int64))
    tables = _lib.TileTables()
    tables.maxId = maxId
    tables.countNew = int(numbered.sum())
    tables.numPairs = len(pairKeys)
    return (tables, rank, flags, pairKeys.astype(numpy.uint64), pairCounts.astype(numpy.uint32))


def three_phase_stitch(tileSegs, tileInfo, nCols, nRows, overlap, simple=False):
    out = numpy.zeros((nRows, nCols), dtype=numpy.uint32)
    luts = {}
    offset = 0
    for (col, row) in sorted(tileInfo.tiles.keys(), key=lambda cr: (cr[1], cr[0])):
        (xpos, ypos, xsize, ysize) = tileInfo.getTile(col, row)
        tile = tileSegs[(col, row)]
        (top, bottom, left, right) = tiling.tileMargins(tileInfo, col, row, xsize, ysize, overlap)
        topB = leftB = None
        if not simple:
            if row > 0:     # LOCAL labels of the neighbours: the resolve maps them to final ids
                topB = tileSegs[(col, row - 1)][-overlap:, :]
            if col > 0:
                leftB = tileSegs[(col - 1, row)][:, -overlap:]
        (tables, rank, flags, pk, pc) = numpy_tile_tables(tile, overlap, topB, leftB, top, bottom,
            left, right)
        (lut, trimmedMax) = tiling.resolveTile(tables, rank, flags, pk, pc, offset,
            luts.get((col, row - 1)), luts.get((col - 1, row)), simple)
        luts[(col, row)] = lut
        win = lut[tile[top:bottom, left:right]]
        out[ypos + top:ypos + bottom, xpos + left:xpos + right] = win
        assert trimmedMax == int(win.max())
        offset = max(offset, trimmedMax)
    return (out, offset)


@pytest.mark.parametrize('name', goldenutil.tiled_names())
def test_three_phase_stitch_matches_reference(name):
    c = goldenutil.load(name)
    m = c['meta']
    img = c['img']
    km = goldenutil.Centres(c['centres'])
    (nB, nR, nC) = img.shape
    ti = tiling.getTilesForFile((nC, nR), m['tileSize'], m['overlapSize'])
    segs = {}
    for ((col, row), (x, y, xs, ys)) in ti.tiles.items():
        sub = numpy.ascontiguousarray(img[:, y:y + ys, x:x + xs])
        segs[(col, row)] = oracle.doShepherdSegmentation(sub, minSegmentSize=m['minSegmentSize'],
            imgNullVal=m['imgNullVal'], fourConnected=m['fourConnected'], kmeansObj=km).segimg
    (mosaic, maxSegId) = three_phase_stitch(segs, ti, nC, nR, m['overlapSize'], m['simpleTileRecode'])
    assert maxSegId == m['maxSegId']
    assert numpy.array_equal(mosaic, c['mosaic'])


def test_three_phase_stitch_random_labels():
    """random blobby label tiles with nulls: exercises ties in the votes and votes for 0"""
    rng = numpy.random.default_rng(11)
    (nR, nC, tileSize, overlap) = (260, 300, 96, 32)
    ti = tiling.getTilesForFile((nC, nR), tileSize, overlap)
    segs = {}
    for ((col, row), (x, y, xs, ys)) in ti.tiles.items():
        coarse = rng.integers(0, 40, (ys // 6 + 2, xs // 6 + 2))
        lab = numpy.kron(coarse, numpy.ones((6, 6), dtype=numpy.int64))[:ys, :xs]
        # make ids contiguous 1..n, keep some null
        (_, inv) = numpy.unique(lab, return_inverse=True)
        lab = inv.reshape(ys, xs).astype(numpy.uint32)
        segs[(col, row)] = lab
    (want, wantMax, _) = oracle.stitchTiles(segs, ti, nC, nR, overlap)
    (got, gotMax) = three_phase_stitch(segs, ti, nC, nR, overlap)
    assert gotMax == wantMax
    assert numpy.array_equal(got, want)


def test_raster_upload_cells_tile_the_tiles():
    """the host-to-device upload cuts the raster along every tile edge; the cells of a tile must
    cover exactly its rectangle, and a cell is sent once however many tiles share it"""
    for (xs, ys, tileSize, overlap, yoff) in ((2745, 2745, 1024, 256, 0), (5000, 3100, 1024, 256, 0),
            (4000, 1800, 1024, 256, 768)):
        ti = tiling.getTilesForFile((xs, ys + yoff), tileSize, overlap)
        order = sorted(ti.tiles.keys(), key=lambda cr: (cr[1], cr[0]))
        mine = [cr for cr in order if ti.tiles[cr][1] >= yoff]      # a rank's row band
        bandRows = max(ti.tiles[cr][1] + ti.tiles[cr][3] for cr in mine) - yoff
        tiles = [(cr,) + tuple(ti.tiles[cr]) for cr in mine]
        up = tiling._RasterUploader(None, 0, [1, 2], bandRows, xs, 2, 0, tiles, yoff)
        sent = numpy.zeros((bandRows, xs), dtype=numpy.int32)
        done = set()
        for t in up.tiles:
            cover = numpy.zeros((bandRows, xs), dtype=bool)
            for cell in up._cellsOf(t):
                (x0, y0, x1, y1) = cell
                assert not cover[y0:y1, x0:x1].any()
                cover[y0:y1, x0:x1] = True
                if cell not in done:
                    done.add(cell)
                    sent[y0:y1, x0:x1] += 1
            (k, x, y, w, h) = t
            want = numpy.zeros((bandRows, xs), dtype=bool)
            want[y:y + h, x:x + w] = True
            assert numpy.array_equal(cover, want)
        assert sent.max() == 1
