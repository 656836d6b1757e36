"""
ctypes front-end of the CPU parity oracle (oracle/shepseg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under pyshepseg_b200/ imports this.

The functions carry the names and argument meaning of the reference functions they
restate (pyshepseg/shepseg.py and pyshepseg/tiling.py) so parity tests read like
calls into the reference.
"""
import ctypes
import os
import subprocess

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBNAME = os.path.join(_HERE, 'liboracle.so')

SegIdType = numpy.uint32
SEGNULLVAL = 0
MINSEGID = 1

_DTYPE_CODES = {
    numpy.dtype(numpy.uint8): 0,
    numpy.dtype(numpy.uint16): 1,
    numpy.dtype(numpy.int16): 2,
    numpy.dtype(numpy.uint32): 3,
    numpy.dtype(numpy.int32): 4,
}


def build():
    """Compile liboracle.so next to this file (make; gcc)."""
    subprocess.check_call(['make', '-s', '-C', _HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, 'shepseg_oracle.c')
        if (not os.path.exists(_LIBNAME) or
                os.path.getmtime(_LIBNAME) < os.path.getmtime(src)):
            build()
        _lib = ctypes.CDLL(_LIBNAME)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError('oracle %s failed with code %d' % (what, rc))


def _img(img):
    img = numpy.ascontiguousarray(img)
    if img.dtype not in _DTYPE_CODES:
        raise TypeError('oracle: unsupported image dtype %s' % img.dtype)
    return img, _DTYPE_CODES[img.dtype]


def centresOf(kmeansObj):
    """Accepts a fitted sklearn KMeans or a bare (k, nBands) array."""
    c = getattr(kmeansObj, 'cluster_centers_', kmeansObj)
    return numpy.ascontiguousarray(c, dtype=numpy.float64)


def applySpectralClusters(kmeansObj, img, imgNullVal):
    """shepseg.py:317-361"""
    img, dt = _img(img)
    (nBands, nRows, nCols) = img.shape
    centres = centresOf(kmeansObj)
    out = numpy.empty((nRows, nCols), dtype=numpy.int32)
    rc = lib().orc_assign(_ptr(img), dt, nBands, ctypes.c_int64(nRows), ctypes.c_int64(nCols),
        _ptr(centres), centres.shape[0], int(imgNullVal is not None),
        ctypes.c_double(0.0 if imgNullVal is None else float(imgNullVal)), _ptr(out))
    _check(rc, 'assign')
    return out


def clump(img, ignoreVal, fourConnected=True, clumpId=1):
    """shepseg.py:452-541"""
    img = numpy.ascontiguousarray(img, dtype=numpy.int32)
    (nRows, nCols) = img.shape
    out = numpy.empty((nRows, nCols), dtype=SegIdType)
    nextId = ctypes.c_uint32(0)
    rc = lib().orc_clump(_ptr(img), ctypes.c_int64(nRows), ctypes.c_int64(nCols),
        ctypes.c_int32(int(ignoreVal)), int(bool(fourConnected)), ctypes.c_uint32(clumpId),
        _ptr(out), ctypes.byref(nextId))
    _check(rc, 'clump')
    return (out, int(nextId.value))


def makeSegSize(seg):
    """shepseg.py:544-569"""
    seg = numpy.ascontiguousarray(seg, dtype=SegIdType)
    n = int(seg.max()) + 1 if seg.size else 1
    segSize = numpy.zeros(n, dtype=numpy.uint32)
    rc = lib().orc_make_seg_size(_ptr(seg), ctypes.c_int64(seg.size), _ptr(segSize),
        ctypes.c_int64(n))
    _check(rc, 'makeSegSize')
    return segSize


def relabelSegments(seg, segSize, minSegId):
    """shepseg.py:739-777 (in place)"""
    assert seg.dtype == SegIdType and seg.flags.c_contiguous
    segSize = numpy.ascontiguousarray(segSize, dtype=numpy.uint32)
    rc = lib().orc_relabel_segments(_ptr(seg), ctypes.c_int64(seg.size), _ptr(segSize),
        ctypes.c_int64(len(segSize)), ctypes.c_uint32(minSegId))
    _check(rc, 'relabelSegments')


def eliminateSinglePixels(img, seg, segSize, minSegId, maxSegId, fourConnected):
    """shepseg.py:572-615 (seg and segSize modified in place)"""
    img, dt = _img(img)
    (nBands, nRows, nCols) = img.shape
    assert seg.dtype == SegIdType and seg.flags.c_contiguous
    assert segSize.dtype == numpy.uint32 and segSize.flags.c_contiguous
    moved = ctypes.c_int64(0)
    rc = lib().orc_eliminate_single_pixels(_ptr(img), dt, nBands, ctypes.c_int64(nRows),
        ctypes.c_int64(nCols), _ptr(seg), _ptr(segSize), ctypes.c_int64(len(segSize)),
        ctypes.c_uint32(minSegId), int(bool(fourConnected)), ctypes.byref(moved))
    _check(rc, 'eliminateSinglePixels')
    return int(moved.value)


def spectralThreshold(maxSpectralDiff):
    """maxSpectralDiff**2 evaluated in maxSpectralDiff's own type as numba does
    (shepseg.py:1060): float32 product for numpy.float32, float64 otherwise."""
    if isinstance(maxSpectralDiff, numpy.float32):
        return float(numpy.float32(maxSpectralDiff) * numpy.float32(maxSpectralDiff))
    if isinstance(maxSpectralDiff, (int, numpy.integer)):
        return float(int(maxSpectralDiff) ** 2)
    return float(maxSpectralDiff) ** 2


def eliminateSmallSegments(seg, img, maxSegId, minSegSize, maxSpectralDiff,
        fourConnected, minSegId):
    """shepseg.py:918-1000 (seg modified in place); returns number eliminated"""
    img, dt = _img(img)
    (nBands, nRows, nCols) = img.shape
    assert seg.dtype == SegIdType and seg.flags.c_contiguous
    numElim = ctypes.c_int64(0)
    rc = lib().orc_eliminate_small_segments(_ptr(seg), _ptr(img), dt, nBands,
        ctypes.c_int64(nRows), ctypes.c_int64(nCols), ctypes.c_uint32(int(maxSegId)),
        int(minSegSize), ctypes.c_double(spectralThreshold(maxSpectralDiff)),
        int(bool(fourConnected)), ctypes.c_uint32(minSegId), ctypes.byref(numElim))
    _check(rc, 'eliminateSmallSegments')
    return int(numElim.value)


def autoMaxSpectralDiff(km, maxSpectralDiff, distPcntile):
    """shepseg.py:400-449.  Host-side numpy, same expressions as the reference so
    that the float32 results are the same numbers."""
    centres = centresOf(km)
    numClusters = centres.shape[0]
    numPairs = numClusters * (numClusters - 1) // 2
    clusterDist = numpy.full(numPairs, -1, dtype=numpy.float32)
    k = 0
    for i in range(numClusters - 1):
        for j in range(i + 1, numClusters):
            clusterDist[k] = numpy.sqrt(((centres[i] - centres[j])**2).sum())
            k += 1
    if isinstance(maxSpectralDiff, str) and maxSpectralDiff == 'auto':
        maxSpectralDiff = numpy.percentile(clusterDist, distPcntile)
    elif maxSpectralDiff is None:
        maxSpectralDiff = 10 * clusterDist.max()
    return maxSpectralDiff


class SegmentationResult(object):
    """shepseg.py:104-127"""
    def __init__(self):
        self.segimg = None
        self.kmeans = None
        self.maxSpectralDiff = None
        self.singlePixelsEliminated = None
        self.smallSegmentsEliminated = None
        self.numClumps = None


def doShepherdSegmentation(img, numClusters=60, clusterSubsamplePcnt=1,
        minSegmentSize=50, maxSpectralDiff='auto', imgNullVal=None,
        fourConnected=True, verbose=False, fixedKMeansInit=False,
        kmeansObj=None, spectDistPcntile=50):
    """shepseg.py:130-249 with kmeansObj required (the fit is not part of the path)."""
    if kmeansObj is None:
        raise ValueError('oracle.doShepherdSegmentation needs kmeansObj')
    img, dt = _img(img)
    (nBands, nRows, nCols) = img.shape
    centres = centresOf(kmeansObj)
    msd = autoMaxSpectralDiff(kmeansObj, maxSpectralDiff, spectDistPcntile)
    seg = numpy.empty((nRows, nCols), dtype=SegIdType)
    nSeg = ctypes.c_uint32(0)
    nSingle = ctypes.c_uint32(0)
    nSmall = ctypes.c_int64(0)
    nClumps = ctypes.c_uint32(0)
    rc = lib().orc_segment(_ptr(img), dt, nBands, ctypes.c_int64(nRows), ctypes.c_int64(nCols),
        _ptr(centres), centres.shape[0], int(imgNullVal is not None),
        ctypes.c_double(0.0 if imgNullVal is None else float(imgNullVal)),
        int(bool(fourConnected)), int(minSegmentSize), ctypes.c_double(spectralThreshold(msd)),
        _ptr(seg), ctypes.byref(nSeg), ctypes.byref(nSingle), ctypes.byref(nSmall),
        ctypes.byref(nClumps))
    _check(rc, 'segment')
    res = SegmentationResult()
    res.segimg = seg
    res.kmeans = kmeansObj
    res.maxSpectralDiff = msd
    res.singlePixelsEliminated = numpy.uint32(nSingle.value)
    res.smallSegmentsEliminated = int(nSmall.value)
    res.numClumps = int(nClumps.value)
    return res


# ---------------------------------------------------------------------------------------
# Tiled path: tile layout and stitching (tiling.py:376-443, 950-1306)
# ---------------------------------------------------------------------------------------
class TileInfo(object):
    """tiling.py:317-373"""
    def __init__(self):
        self.tiles = {}
        self.ncols = None
        self.nrows = None

    def addTile(self, xpos, ypos, xsize, ysize, col, row):
        self.tiles[(col, row)] = (xpos, ypos, xsize, ysize)

    def getNumTiles(self):
        return len(self.tiles)

    def getTile(self, col, row):
        return self.tiles[(col, row)]


def _axisTiles(rasterSize, tileSize, overlapSize):
    """One axis of tiling.py:405-438: (pos, size) of every tile along it.  Tiles step by
    tileSize-overlapSize; the tile after which another whole tile would not fit grows to
    the raster edge."""
    spans = []
    pos = 0
    while True:
        size = tileSize
        last = (pos + size * 2) > rasterSize
        if last:
            size = rasterSize - pos
        if size > 0:
            spans.append((pos, size))
        if last:
            return spans
        pos += tileSize - overlapSize


def getTilesForFile(rasterXSize, rasterYSize, tileSize, overlapSize):
    """tiling.py:376-443, taking the raster size instead of a GDAL dataset.  The grid is
    separable, so it is the product of the two axis layouts."""
    tileInfo = TileInfo()
    xs = _axisTiles(int(rasterXSize), int(tileSize), int(overlapSize))
    ys = _axisTiles(int(rasterYSize), int(tileSize), int(overlapSize))
    for (row, (ypos, ysize)) in enumerate(ys):
        for (col, (xpos, xsize)) in enumerate(xs):
            tileInfo.addTile(xpos, ypos, xsize, ysize, col, row)
    tileInfo.ncols = len(xs)
    tileInfo.nrows = len(ys)
    return tileInfo


def recodeTile(tileData, maxSegId, overlapSize, topOverlapB, leftOverlapB,
        top, bottom, left, right):
    """tiling.py:1066-1126 (+1128-1306).  Returns (newTileData, newMaxSegId)."""
    tileData = numpy.ascontiguousarray(tileData, dtype=SegIdType)
    (ysize, xsize) = tileData.shape
    out = numpy.empty_like(tileData)
    newMax = ctypes.c_uint32(0)
    tb = None if topOverlapB is None else numpy.ascontiguousarray(topOverlapB, dtype=SegIdType)
    lb = None if leftOverlapB is None else numpy.ascontiguousarray(leftOverlapB, dtype=SegIdType)
    rc = lib().orc_recode_tile(_ptr(tileData), ctypes.c_int64(ysize), ctypes.c_int64(xsize),
        ctypes.c_int64(overlapSize), None if tb is None else _ptr(tb),
        None if lb is None else _ptr(lb), ctypes.c_uint32(int(maxSegId)),
        ctypes.c_int64(top), ctypes.c_int64(bottom), ctypes.c_int64(left), ctypes.c_int64(right),
        _ptr(out), ctypes.byref(newMax))
    _check(rc, 'recodeTile')
    return (out, int(newMax.value))


def stitchTiles(tileSegs, tileInfo, rasterXSize, rasterYSize, overlapSize,
        simpleTileRecode=False):
    """
    tiling.py:950-1064 without the file I/O: tileSegs maps (col, row) -> uint32 label
    array of that tile.  Returns (mosaic, maxSegId, histogram).
    """
    marginSize = int(overlapSize / 2)
    out = numpy.zeros((rasterYSize, rasterXSize), dtype=SegIdType)
    colRowList = sorted(tileInfo.tiles.keys(), key=lambda x: (x[1], x[0]))
    maxSegId = 0
    overlapCache = {}
    for (col, row) in colRowList:
        (xpos, ypos, xsize, ysize) = tileInfo.getTile(col, row)
        tileData = numpy.array(tileSegs[(col, row)], dtype=SegIdType)
        top = marginSize
        bottom = ysize - marginSize
        left = marginSize
        right = xsize - marginSize
        xout = xpos + marginSize
        yout = ypos + marginSize
        haveRight = True
        haveBottom = True
        if row == 0:
            top = 0
            yout = ypos
        if row == (tileInfo.nrows - 1):
            bottom = ysize
            haveBottom = False
        if col == 0:
            left = 0
            xout = xpos
        if col == (tileInfo.ncols - 1):
            right = xsize
            haveRight = False
        if simpleTileRecode:
            nullmask = (tileData == SEGNULLVAL)
            tileData += SegIdType(maxSegId)
            tileData[nullmask] = SEGNULLVAL
        else:
            topB = overlapCache.get(('bottom', col, row - 1)) if row > 0 else None
            leftB = overlapCache.get(('right', col - 1, row)) if col > 0 else None
            (tileData, _) = recodeTile(tileData, maxSegId, overlapSize, topB, leftB,
                top, bottom, left, right)
        trimmed = tileData[top:bottom, left:right]
        out[yout:yout + trimmed.shape[0], xout:xout + trimmed.shape[1]] = trimmed
        if haveRight:
            overlapCache[('right', col, row)] = tileData[:, -overlapSize:].copy()
        if haveBottom:
            overlapCache[('bottom', col, row)] = tileData[-overlapSize:, :].copy()
        maxSegId = max(maxSegId, int(trimmed.max()))
    hist = numpy.bincount(out.ravel()).astype(numpy.float64) if out.size else numpy.zeros(1)
    hist[0] = 0
    return (out, maxSegId, hist)
