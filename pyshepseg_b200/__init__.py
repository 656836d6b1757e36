"""
pyshepseg_b200 -- the Shepherd segmentation hot path of ubarsc/pyshepseg on NVIDIA B200.

Modules mirror the reference's: `shepseg` (in-memory segmentation of one tile) and `tiling`
(tiled segmentation of a raster file with on-device stitching).  `_lib` is the ctypes
binding of the C ABI (include/shepseg_b200.h); `synth` generates the synthetic rasters used
by the tests and the benchmark.
"""
__version__ = '0.1.0'
