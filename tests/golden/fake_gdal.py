"""
A numpy-backed stand-in for osgeo.gdal / osgeo.gdal_array, just large enough that the
UNMODIFIED reference pyshepseg.tiling runs end to end on in-memory rasters.

Only used by tests/golden/make_golden.py (in the development container, where the
reference is mounted and GDAL is absent) to produce tiled golden fixtures.  "Files" are
entries of the REGISTRY dict keyed by filename.
"""
import sys
import types

import numpy

REGISTRY = {}

GDT_Byte, GDT_UInt16, GDT_Int16, GDT_UInt32, GDT_Int32, GDT_Float32, GDT_Float64 = 1, 2, 3, 4, 5, 6, 7
_NP2GDT = {numpy.dtype(numpy.uint8): GDT_Byte, numpy.dtype(numpy.uint16): GDT_UInt16,
    numpy.dtype(numpy.int16): GDT_Int16, numpy.dtype(numpy.uint32): GDT_UInt32,
    numpy.dtype(numpy.int32): GDT_Int32, numpy.dtype(numpy.float32): GDT_Float32,
    numpy.dtype(numpy.float64): GDT_Float64}
_GDT2NP = dict((v, k) for (k, v) in _NP2GDT.items())

GFU_Generic, GFU_PixelCount, GFU_Red, GFU_Green, GFU_Blue, GFU_Alpha = 0, 1, 6, 7, 8, 9
GFT_Integer, GFT_Real, GFT_String = 0, 1, 2
GA_ReadOnly, GA_Update = 0, 1


class FakeRAT(object):
    def __init__(self):
        self.cols = []      # (name, type, usage, array)
        self.nrows = 0

    def GetRowCount(self):
        return self.nrows

    def SetRowCount(self, n):
        self.nrows = n

    def GetColumnCount(self):
        return len(self.cols)

    def GetColOfUsage(self, usage):
        for (i, c) in enumerate(self.cols):
            if c[2] == usage:
                return i
        return -1

    def GetNameOfCol(self, i):
        return self.cols[i][0]

    def CreateColumn(self, name, typ, usage):
        self.cols.append([name, typ, usage, None])

    def WriteArray(self, arr, col, start=0):
        self.cols[col][3] = numpy.array(arr)

    def ReadAsArray(self, col, start=0, length=None):
        return self.cols[col][3]


class FakeBand(object):
    def __init__(self, arr):
        self.arr = arr
        self.nodata = None
        self.meta = {}
        self.rat = FakeRAT()
        self.overviews = []
        self.DataType = _NP2GDT[arr.dtype]

    @property
    def XSize(self):
        return self.arr.shape[1]

    @property
    def YSize(self):
        return self.arr.shape[0]

    def ReadAsArray(self, xoff=0, yoff=0, xsize=None, ysize=None):
        if xsize is None:
            xsize = self.XSize - xoff
        if ysize is None:
            ysize = self.YSize - yoff
        return self.arr[yoff:yoff + ysize, xoff:xoff + xsize].copy()

    def WriteArray(self, a, xoff=0, yoff=0):
        self.arr[yoff:yoff + a.shape[0], xoff:xoff + a.shape[1]] = a

    def GetNoDataValue(self):
        return self.nodata

    def SetNoDataValue(self, v):
        self.nodata = v

    def SetMetadataItem(self, k, v):
        self.meta[k] = v

    def GetMetadataItem(self, k):
        return self.meta.get(k)

    def GetDefaultRAT(self):
        return self.rat

    def GetOverview(self, i):
        return self.overviews[i]

    def GetOverviewCount(self):
        return len(self.overviews)

    def FlushCache(self):
        pass


class FakeDataset(object):
    def __init__(self, bands):
        self.bands = [FakeBand(b) for b in bands]
        self.proj = ''
        self.gt = (0.0, 1.0, 0.0, 0.0, 0.0, -1.0)

    @property
    def RasterXSize(self):
        return self.bands[0].XSize

    @property
    def RasterYSize(self):
        return self.bands[0].YSize

    @property
    def RasterCount(self):
        return len(self.bands)

    def GetRasterBand(self, i):
        return self.bands[i - 1]

    def GetProjection(self):
        return self.proj

    def SetProjection(self, p):
        self.proj = p

    def GetGeoTransform(self):
        return self.gt

    def SetGeoTransform(self, gt):
        self.gt = tuple(gt)

    def BuildOverviews(self, resampling, levels):
        for b in self.bands:
            b.overviews = []
            for lvl in levels:
                shape = ((b.YSize + lvl - 1) // lvl, (b.XSize + lvl - 1) // lvl)
                b.overviews.append(FakeBand(numpy.zeros(shape, dtype=b.arr.dtype)))

    def ReadAsArray(self):
        if len(self.bands) == 1:
            return self.bands[0].arr.copy()
        return numpy.array([b.arr for b in self.bands])

    def FlushCache(self):
        pass


class FakeDriver(object):
    def __init__(self, name):
        self.name = name

    def Create(self, filename, xsize, ysize, nbands, gdaltype, options=None):
        ds = FakeDataset([numpy.zeros((ysize, xsize), dtype=_GDT2NP[gdaltype])
            for _ in range(nbands)])
        REGISTRY[filename] = ds
        return ds

    def Delete(self, filename):
        REGISTRY.pop(filename, None)

    def GetMetadataItem(self, key):
        if key == 'DMD_EXTENSION':
            return 'kea'
        return None


def put_image(filename, img, nodata=None):
    """Register a (nBands, nRows, nCols) array as a readable fake file."""
    ds = FakeDataset([numpy.ascontiguousarray(b) for b in img])
    for b in ds.bands:
        b.nodata = nodata
    REGISTRY[filename] = ds
    return ds


def install():
    """Place fake osgeo, osgeo.gdal, osgeo.gdal_array, osgeo.osr into sys.modules."""
    osgeo = types.ModuleType('osgeo')
    gdal = types.ModuleType('osgeo.gdal')
    gdal_array = types.ModuleType('osgeo.gdal_array')
    osr = types.ModuleType('osgeo.osr')
    for (k, v) in list(globals().items()):
        if k.startswith(('GDT_', 'GFU_', 'GFT_', 'GA_')):
            setattr(gdal, k, v)
    gdal.UseExceptions = lambda: None
    gdal.Dataset = FakeDataset
    gdal.Band = FakeBand
    gdal.Open = lambda filename, mode=0: REGISTRY[filename]
    gdal.GetDriverByName = lambda name: FakeDriver(name)
    gdal.IdentifyDriver = lambda filename: FakeDriver('FAKE')
    gdal_array.NumericTypeCodeToGDALTypeCode = lambda t: _NP2GDT[numpy.dtype(t)]
    osgeo.gdal = gdal
    osgeo.gdal_array = gdal_array
    osgeo.osr = osr
    sys.modules['osgeo'] = osgeo
    sys.modules['osgeo.gdal'] = gdal
    sys.modules['osgeo.gdal_array'] = gdal_array
    sys.modules['osgeo.osr'] = osr
    return gdal
