"""
Raster input / output for the tiled driver.

The reference reads and writes through GDAL (tiling.py:774, 961-975, 1436-1443).  GDAL is
used here too when `osgeo` is importable.  Where it is not (the build and GPU images have no
GDAL) a small built-in reader / writer covers what the path needs:

  * uncompressed, strip-organised TIFF and BigTIFF, 8 / 16 / 32 bit integer samples, pixel- or
    band-interleaved (read); band-interleaved, one strip per band (write) -- driver 'GTiff';
  * numpy .npy arrays of shape (nBands, nRows, nCols) or (nRows, nCols)            -- driver 'NPY';
  * in-memory arrays (MemoryRaster), mainly for tests.

Window reads return band-sequential (nBands, ysize, xsize) arrays, which is the layout the
segmentation takes (shepseg.py:140).
"""
import json
import mmap
import os
import struct

import numpy

try:
    from osgeo import gdal as _gdal
    _gdal.UseExceptions()
except Exception:      # pragma: no cover - GDAL is absent in the build image
    _gdal = None


class RasterError(Exception):
    pass


def haveGDAL():
    return _gdal is not None


# ---------------------------------------------------------------------------------------
# Sources
# ---------------------------------------------------------------------------------------
class RasterSource(object):
    """xsize, ysize, count, dtype, nodata (one per band or None), projection, geotransform."""
    projection = ''
    geotransform = (0.0, 1.0, 0.0, 0.0, 0.0, -1.0)

    def readWindow(self, bandNumbers, xoff, yoff, xsize, ysize, out=None):
        raise NotImplementedError

    def close(self):
        pass


class MemoryRaster(RasterSource):
    """A (count, ysize, xsize) array as a raster.  With `yoff` / `fullYsize` (and `xoff` /
    `fullXsize`) the array is the window [yoff, yoff + ysize) x [xoff, xoff + xsize) of a larger
    raster of fullYsize x fullXsize: what one rank of a sharded run holds of the mosaic (window
    reads are then relative to it)."""
    def __init__(self, img, nodata=None, yoff=0, fullYsize=None, xoff=0, fullXsize=None):
        img = numpy.asarray(img)
        if img.ndim == 2:
            img = img[None]
        self.img = img
        (self.count, self.ysize, self.xsize) = img.shape
        self.dtype = img.dtype
        self.nodata = [nodata] * self.count
        self.yoff = int(yoff)
        self.fullYsize = self.ysize if fullYsize is None else int(fullYsize)
        self.xoff = int(xoff)
        self.fullXsize = self.xsize if fullXsize is None else int(fullXsize)

    def readWindow(self, bandNumbers, xoff, yoff, xsize, ysize, out=None):
        if out is None:
            out = numpy.empty((len(bandNumbers), ysize, xsize), dtype=self.dtype)
        for (i, b) in enumerate(bandNumbers):
            out[i] = self.img[b - 1, yoff:yoff + ysize, xoff:xoff + xsize]
        return out


class NpySource(MemoryRaster):
    def __init__(self, filename):
        img = numpy.load(filename, mmap_mode='r')
        nodata = None
        side = filename + '.json'
        if os.path.exists(side):
            meta = json.load(open(side))
            nodata = meta.get('nodata')
        MemoryRaster.__init__(self, img, nodata)


_TIFF_TYPES = {1: ('B', 1), 2: ('c', 1), 3: ('H', 2), 4: ('I', 4), 5: ('II', 8), 6: ('b', 1),
    8: ('h', 2), 9: ('i', 4), 11: ('f', 4), 12: ('d', 8), 16: ('Q', 8), 17: ('q', 8), 18: ('Q', 8)}


def _readIFD(f):
    """Parse the first IFD of a little/big-endian TIFF or BigTIFF: {tag: tuple of values}."""
    head = f.read(16)
    if head[:2] == b'II':
        e = '<'
    elif head[:2] == b'MM':
        e = '>'
    else:
        raise RasterError('not a TIFF file')
    magic = struct.unpack(e + 'H', head[2:4])[0]
    if magic == 42:
        (big, ifdOff) = (False, struct.unpack(e + 'I', head[4:8])[0])
    elif magic == 43:
        (big, ifdOff) = (True, struct.unpack(e + 'Q', head[8:16])[0])
    else:
        raise RasterError('not a TIFF file')
    f.seek(ifdOff)
    n = struct.unpack(e + ('Q' if big else 'H'), f.read(8 if big else 2))[0]
    entrySize = 20 if big else 12
    raw = f.read(n * entrySize)
    tags = {}
    for i in range(n):
        ent = raw[i * entrySize:(i + 1) * entrySize]
        (tag, typ) = struct.unpack(e + 'HH', ent[:4])
        cnt = struct.unpack(e + ('Q' if big else 'I'), ent[4:(12 if big else 8)])[0]
        valField = ent[(12 if big else 8):]
        if typ not in _TIFF_TYPES:
            continue
        (code, size) = _TIFF_TYPES[typ]
        total = size * cnt
        if total <= len(valField):
            data = valField[:total]
        else:
            off = struct.unpack(e + ('Q' if big else 'I'), valField)[0]
            pos = f.tell()
            f.seek(off)
            data = f.read(total)
            f.seek(pos)
        if typ == 2:
            tags[tag] = data.rstrip(b'\x00').decode('latin1')
        elif typ == 5:
            vals = struct.unpack(e + 'I' * (2 * cnt), data)
            tags[tag] = tuple(vals[2 * j] / max(vals[2 * j + 1], 1) for j in range(cnt))
        else:
            tags[tag] = struct.unpack(e + code * cnt, data)
    return (e, tags)


class TiffSource(RasterSource):
    """Uncompressed strip TIFF / BigTIFF whose strips are laid out contiguously."""
    def __init__(self, filename):
        self.f = open(filename, 'rb')
        (e, t) = _readIFD(self.f)
        if t.get(259, (1,))[0] != 1:
            raise RasterError('compressed TIFF needs GDAL, which is not available here')
        if 322 in t or 324 in t:
            raise RasterError('tiled TIFF needs GDAL, which is not available here')
        self.xsize = int(t[256][0])
        self.ysize = int(t[257][0])
        self.count = int(t.get(277, (1,))[0])
        bits = t.get(258, (1,))
        fmt = t.get(339, (1,))[0]
        if len(set(bits)) != 1:
            raise RasterError('bands of different bit depth are not supported')
        kind = {1: 'u', 2: 'i'}.get(fmt)
        if kind is None or bits[0] not in (8, 16, 32):
            raise RasterError('only 8/16/32 bit integer TIFF samples are supported')
        self.dtype = numpy.dtype(e + kind + str(bits[0] // 8)) if bits[0] > 8 else numpy.dtype(kind + '1')
        planar = t.get(284, (1,))[0]
        offsets = t[273]
        counts = t[279]
        start = offsets[0]
        pos = start
        for (o, c) in zip(offsets, counts):
            if o != pos:
                raise RasterError('TIFF strips are not contiguous; needs GDAL')
            pos += c
        nbytes = self.xsize * self.ysize * self.count * self.dtype.itemsize
        if pos - start != nbytes:
            raise RasterError('unexpected TIFF data size')
        self.mm = mmap.mmap(self.f.fileno(), 0, access=mmap.ACCESS_READ)
        if planar == 2 or self.count == 1:
            self.data = numpy.frombuffer(self.mm, dtype=self.dtype, count=nbytes // self.dtype.itemsize,
                offset=start).reshape(self.count, self.ysize, self.xsize)
            self.pixelInterleaved = False
        else:
            self.data = numpy.frombuffer(self.mm, dtype=self.dtype, count=nbytes // self.dtype.itemsize,
                offset=start).reshape(self.ysize, self.xsize, self.count)
            self.pixelInterleaved = True
        nd = t.get(42113)
        nodata = None
        if nd is not None:
            try:
                nodata = float(nd)
                if nodata == int(nodata):
                    nodata = int(nodata)
            except ValueError:
                nodata = None
        self.nodata = [nodata] * self.count
        gt = (0.0, 1.0, 0.0, 0.0, 0.0, -1.0)
        if 33550 in t and 33922 in t:      # ModelPixelScale + ModelTiepoint
            (sx, sy) = t[33550][:2]
            tp = t[33922]
            gt = (tp[3] - tp[0] * sx, sx, 0.0, tp[4] + tp[1] * sy, 0.0, -sy)
        self.geotransform = gt
        self.geokeys = dict((k, t[k]) for k in (33550, 33922, 34735, 34736, 34737) if k in t)

    def readWindow(self, bandNumbers, xoff, yoff, xsize, ysize, out=None):
        if out is None:
            out = numpy.empty((len(bandNumbers), ysize, xsize), dtype=self.dtype.newbyteorder('='))
        for (i, b) in enumerate(bandNumbers):
            if self.pixelInterleaved:
                out[i] = self.data[yoff:yoff + ysize, xoff:xoff + xsize, b - 1]
            else:
                out[i] = self.data[b - 1, yoff:yoff + ysize, xoff:xoff + xsize]
        return out

    def close(self):
        self.data = None
        try:
            self.mm.close()
        except BufferError:
            pass
        self.f.close()


class GdalSource(RasterSource):      # pragma: no cover - exercised only where GDAL exists
    def __init__(self, filename):
        self.ds = _gdal.Open(filename)
        self.xsize = self.ds.RasterXSize
        self.ysize = self.ds.RasterYSize
        self.count = self.ds.RasterCount
        from osgeo import gdal_array
        self.dtype = numpy.dtype(gdal_array.GDALTypeCodeToNumericTypeCode(
            self.ds.GetRasterBand(1).DataType))
        self.nodata = [self.ds.GetRasterBand(i + 1).GetNoDataValue() for i in range(self.count)]
        self.projection = self.ds.GetProjection()
        self.geotransform = self.ds.GetGeoTransform()

    def readWindow(self, bandNumbers, xoff, yoff, xsize, ysize, out=None):
        if out is None:
            out = numpy.empty((len(bandNumbers), ysize, xsize), dtype=self.dtype)
        for (i, b) in enumerate(bandNumbers):
            out[i] = self.ds.GetRasterBand(b).ReadAsArray(xoff, yoff, xsize, ysize)
        return out

    def close(self):
        self.ds = None


def openRaster(infile):
    """A RasterSource for a filename, a numpy array or an existing RasterSource."""
    if isinstance(infile, RasterSource):
        return infile
    if isinstance(infile, numpy.ndarray):
        return MemoryRaster(infile)
    if not os.path.exists(infile):
        raise RasterError("input raster '%s' does not exist" % infile)
    if infile.lower().endswith('.npy'):
        return NpySource(infile)
    if _gdal is not None:
        return GdalSource(infile)
    return TiffSource(infile)


# ---------------------------------------------------------------------------------------
# Sinks
# ---------------------------------------------------------------------------------------
class RasterSink(object):
    """Single-band uint32 output raster written window by window."""
    def write(self, arr, xoff, yoff):
        raise NotImplementedError

    def setNoData(self, value):
        pass

    def setMetadataItem(self, key, value):
        pass

    def writeHistogram(self, hist):
        pass

    def writeOverviews(self, arr, xoff, yoff):
        """All overview levels of the window arr written at (xoff, yoff) (tiling.py:1360-1383)."""
        for (j, lvl) in enumerate(getattr(self, 'levels', None) or []):
            self.writeOverviewLevel(j, lvl, arr[lvl // 2::lvl, lvl // 2::lvl], xoff // lvl, yoff // lvl)

    def writeOverviewLevel(self, j, lvl, sub, xs, ys):
        """sub = the window sub-sampled for level number j (= lvl), to go at (xs, ys) of that level."""
        pass

    def close(self):
        pass


class MemorySink(RasterSink):
    """The output raster as a numpy array.  `array` (optional) is caller-owned memory to write
    into, e.g. pinned host memory; with `yoff` / `xoff` it is the window of a larger mosaic that
    starts at mosaic row yoff and column xoff (one rank of a sharded run).  Overviews are kept per level in
    `overviews` when `levels` is given (tiling.py:1360-1383)."""
    levels = ()
    overviews = {}
    yoff = 0
    xoff = 0

    def __init__(self, xsize, ysize, dtype=numpy.uint32, array=None, yoff=0, levels=None, xoff=0):
        if array is None:
            array = numpy.zeros((ysize, xsize), dtype=dtype)
        elif array.shape != (ysize, xsize):
            raise RasterError('array of shape %s given for a %d x %d raster' % (array.shape, ysize, xsize))
        self.array = array
        self.yoff = int(yoff)
        self.xoff = int(xoff)
        self.metadata = {}
        self.nodata = None
        self.hist = None
        self.levels = list(levels) if levels else []
        self.overviews = dict((lvl, numpy.zeros((-(-ysize // lvl), -(-xsize // lvl)), dtype=dtype))
            for lvl in self.levels)

    def writeOverviewLevel(self, j, lvl, sub, xs, ys):
        ov = self.overviews[lvl]
        sub = sub[:ov.shape[0] - ys, :ov.shape[1] - xs]      # (do not go off the edges)
        ov[ys:ys + sub.shape[0], xs:xs + sub.shape[1]] = sub

    def write(self, arr, xoff, yoff):
        self.array[yoff:yoff + arr.shape[0], xoff:xoff + arr.shape[1]] = arr

    def setNoData(self, value):
        self.nodata = value

    def setMetadataItem(self, key, value):
        self.metadata[key] = value

    def writeHistogram(self, hist):
        self.hist = hist


def _sidecar(filename, nodata, metadata, hist):
    meta = {'nodata': nodata, 'metadata': metadata}
    if hist is not None:
        numpy.save(filename + '.hist.npy', hist)
        meta['histogram'] = os.path.basename(filename) + '.hist.npy'
    json.dump(meta, open(filename + '.json', 'w'), indent=1)


class NpySink(MemorySink):
    def __init__(self, filename, xsize, ysize, dtype=numpy.uint32):
        self.filename = filename
        self.array = numpy.lib.format.open_memmap(filename, mode='w+', dtype=dtype, shape=(ysize, xsize))
        self.metadata = {}
        self.nodata = None
        self.hist = None

    def close(self):
        if self.array is not None:
            self.array.flush()
            self.array = None
            _sidecar(self.filename, self.nodata, self.metadata, self.hist)


class TiffSink(MemorySink):
    """One-band uncompressed TIFF (BigTIFF above 4 GB), data memory-mapped for window writes."""
    def __init__(self, filename, xsize, ysize, dtype=numpy.uint32, nodata=0, source=None):
        self.filename = filename
        dtype = numpy.dtype(dtype)
        nbytes = xsize * ysize * dtype.itemsize
        big = nbytes + 4096 >= 2**32
        fmt = {'u': 1, 'i': 2}[dtype.kind]
        nd = (str(nodata) + '\x00').encode('latin1')
        ents = [(256, 4, [xsize]), (257, 4, [ysize]), (258, 3, [dtype.itemsize * 8]), (259, 3, [1]),
            (262, 3, [1]), (273, 16 if big else 4, [0]), (277, 3, [1]), (278, 4, [ysize]),
            (279, 16 if big else 4, [nbytes]), (339, 3, [fmt]), (42113, 2, nd)]
        if source is not None and getattr(source, 'geokeys', None):
            for (tag, vals) in source.geokeys.items():
                if tag in (33550, 33922, 34736):
                    ents.append((tag, 12, list(vals)))
                elif tag == 34735:
                    ents.append((tag, 3, list(vals)))
                elif tag == 34737:
                    ents.append((tag, 2, (vals + '\x00').encode('latin1')))
        ents.sort(key=lambda x: x[0])
        headSize = 16 if big else 8
        entrySize = 20 if big else 12
        ifdSize = (8 if big else 2) + len(ents) * entrySize + (8 if big else 4)
        extraOff = headSize + ifdSize
        extra = b''
        packed = []
        inl = 8 if big else 4
        for (tag, typ, vals) in ents:
            if typ == 2:
                data = bytes(vals)
                cnt = len(data)
            else:
                (code, size) = _TIFF_TYPES[typ]
                data = struct.pack('<' + code * len(vals), *vals)
                cnt = len(vals)
            if len(data) <= inl:
                field = data + b'\x00' * (inl - len(data))
            else:
                field = struct.pack('<Q' if big else '<I', extraOff + len(extra))
                extra += data + (b'\x00' if len(data) % 2 else b'')
            packed.append((tag, typ, cnt, field))
        dataOff = (extraOff + len(extra) + 15) // 16 * 16
        out = b'II' + (struct.pack('<HHHQ', 43, 8, 0, 16) if big else struct.pack('<HI', 42, 8))
        out += struct.pack('<Q' if big else '<H', len(packed))
        for (tag, typ, cnt, field) in packed:
            if tag == 273:
                field = struct.pack('<Q' if big else '<I', dataOff)
            out += struct.pack('<HH', tag, typ) + struct.pack('<Q' if big else '<I', cnt) + field
        out += struct.pack('<Q' if big else '<I', 0)
        out += extra
        out += b'\x00' * (dataOff - len(out))
        with open(filename, 'wb') as f:
            f.write(out)
            f.truncate(dataOff + nbytes)
        self.array = numpy.memmap(filename, dtype=dtype, mode='r+', offset=dataOff, shape=(ysize, xsize))
        self.metadata = {}
        self.nodata = nodata
        self.hist = None

    def close(self):
        if self.array is not None:
            self.array.flush()
            self.array = None
            _sidecar(self.filename, self.nodata, self.metadata, self.hist)


def overviewLevels(xsize, ysize, finalOutSize=1024):
    """
    The overview levels the reference sets up (setupOverviews, tiling.py:1385-1404): 4, 8, ...
    The reference appends a level and only then re-tests the SAME level before moving on, so it
    always emits one level more than "while the overview is at least 1024 pixels" would
    (10980 pixels -> [4, 8, 16]); that is kept.
    """
    outSize = max(int(xsize), int(ysize))
    levels = []
    i = 2
    ok = (outSize // (2 ** i)) >= finalOutSize
    while ok:
        levels.append(2 ** i)
        ok = (outSize // (2 ** i)) >= finalOutSize
        i += 1
    return levels


class GdalSink(RasterSink):      # pragma: no cover - exercised only where GDAL exists
    """Output through GDAL exactly as the reference sets it up (tiling.py:961-975, 1343-1404)."""
    def __init__(self, filename, xsize, ysize, driver, options, source):
        if os.path.exists(filename):
            _gdal.IdentifyDriver(filename).Delete(filename)
        drvr = _gdal.GetDriverByName(driver)
        self.ds = drvr.Create(filename, xsize, ysize, 1, _gdal.GDT_UInt32, options)
        if source is not None:
            self.ds.SetProjection(source.projection)
            self.ds.SetGeoTransform(source.geotransform)
        self.levels = overviewLevels(xsize, ysize)
        self.ds.BuildOverviews("NEAREST", self.levels)
        self.band = self.ds.GetRasterBand(1)
        self.band.SetMetadataItem('LAYER_TYPE', 'thematic')

    def write(self, arr, xoff, yoff):
        self.band.WriteArray(arr, xoff, yoff)

    def writeOverviewLevel(self, j, lvl, sub, xs, ys):
        ov = self.band.GetOverview(j)
        sub = sub[:ov.YSize - ys, :ov.XSize - xs]
        ov.WriteArray(numpy.ascontiguousarray(sub), xs, ys)

    def setNoData(self, value):
        self.band.SetNoDataValue(value)

    def setMetadataItem(self, key, value):
        self.band.SetMetadataItem(key, value)

    def writeHistogram(self, hist):
        rat = self.band.GetDefaultRAT()
        if rat.GetRowCount() != len(hist):
            rat.SetRowCount(len(hist))
        col = rat.GetColOfUsage(_gdal.GFU_PixelCount)
        if col == -1:
            rat.CreateColumn('Histogram', _gdal.GFT_Real, _gdal.GFU_PixelCount)
            col = rat.GetColumnCount() - 1
        rat.WriteArray(hist, col)

    def close(self):
        if self.ds is not None:
            self.ds.FlushCache()
            self.band = None
            self.ds = None


BUILTIN_DRIVERS = ('GTiff', 'NPY', 'MEM')


def driverAvailable(driver):
    if _gdal is not None:
        return _gdal.GetDriverByName(driver) is not None or driver in ('NPY', 'MEM')
    return driver in BUILTIN_DRIVERS


def createRaster(outfile, xsize, ysize, driver, options, source=None):
    """A RasterSink for the output label raster."""
    if isinstance(outfile, RasterSink):
        return outfile
    if driver == 'MEM' or outfile is None:
        return MemorySink(xsize, ysize, levels=overviewLevels(xsize, ysize))
    if driver == 'NPY':
        return NpySink(outfile, xsize, ysize)
    if _gdal is not None:
        return GdalSink(outfile, xsize, ysize, driver, options, source)
    if driver == 'GTiff':
        return TiffSink(outfile, xsize, ysize, source=source)
    raise RasterError("driver '%s' needs GDAL, which is not available here (built in: %s)" % (
        driver, ', '.join(BUILTIN_DRIVERS)))


def writeImage(filename, img, nodata=None):
    """Write a (nBands, nRows, nCols) integer array as an uncompressed band-interleaved TIFF
    (BigTIFF above 4 GB) -- used to generate the synthetic benchmark rasters locally."""
    img = numpy.asarray(img)
    if img.ndim == 2:
        img = img[None]
    (nB, ys, xs) = img.shape
    dtype = img.dtype
    band = xs * ys * dtype.itemsize
    nbytes = band * nB
    big = nbytes + 4096 >= 2**32
    fmt = {'u': 1, 'i': 2}[dtype.kind]
    lt = 16 if big else 4
    ents = [(256, 4, [xs]), (257, 4, [ys]), (258, 3, [dtype.itemsize * 8] * nB), (259, 3, [1]),
        (262, 3, [1]), (273, lt, [0] * nB), (277, 3, [nB]), (278, 4, [ys]), (279, lt, [band] * nB),
        (284, 3, [2]), (339, 3, [fmt] * nB)]
    if nB > 1:
        ents.append((338, 3, [0] * (nB - 1)))
    if nodata is not None:
        ents.append((42113, 2, (str(nodata) + '\x00').encode('latin1')))
    ents.sort(key=lambda x: x[0])
    headSize = 16 if big else 8
    entrySize = 20 if big else 12
    ifdSize = (8 if big else 2) + len(ents) * entrySize + (8 if big else 4)
    extraOff = headSize + ifdSize
    inl = 8 if big else 4
    # first pass to size the out-of-line area, second to fill the strip offsets
    def pack(dataOff):
        extra = b''
        body = b''
        for (tag, typ, vals) in ents:
            if tag == 273:
                vals = [dataOff + i * band for i in range(nB)]
            if typ == 2:
                data = bytes(vals)
                cnt = len(data)
            else:
                (code, size) = _TIFF_TYPES[typ]
                data = struct.pack('<' + code * len(vals), *vals)
                cnt = len(vals)
            if len(data) <= inl:
                field = data + b'\x00' * (inl - len(data))
            else:
                field = struct.pack('<Q' if big else '<I', extraOff + len(extra))
                extra += data + (b'\x00' if len(data) % 2 else b'')
            body += struct.pack('<HH', tag, typ) + struct.pack('<Q' if big else '<I', cnt) + field
        return (body, extra)
    (body, extra) = pack(0)
    dataOff = (extraOff + len(extra) + 15) // 16 * 16
    (body, extra) = pack(dataOff)
    out = b'II' + (struct.pack('<HHHQ', 43, 8, 0, 16) if big else struct.pack('<HI', 42, 8))
    out += struct.pack('<Q' if big else '<H', len(ents)) + body + struct.pack('<Q' if big else '<I', 0)
    out += extra
    out += b'\x00' * (dataOff - len(out))
    with open(filename, 'wb') as f:
        f.write(out)
        for b in range(nB):
            f.write(numpy.ascontiguousarray(img[b]).data)
