"""
Synthetic multi-band rasters for parity tests and benchmarks.

These are the inputs SURVEY.md section 8(d) / BASELINE.md section 3 name: `synth_v1`
(smooth random field + noise; many small clumps, like real reflectance imagery),
`synth_tiled` (the same statistics but generated block by block so that 10980^2 and
40000^2 rasters never need a float64 copy of the whole band in memory) and `synth_flat`
(Voronoi cells of constant colour, like the image the reference's own functional test
builds in pyshepseg/cmdline/runtests.py:145-265, which drives every clump past the
reference's 10000-pixel cap).
"""
import numpy


def _bilinear_zoom(coarse, cell, nRows, nCols, row0=0, col0=0):
    """Bilinear upsampling of a coarse grid by `cell`, window (row0.., col0..)."""
    r = (numpy.arange(row0, row0 + nRows, dtype=numpy.float64)) / cell
    c = (numpy.arange(col0, col0 + nCols, dtype=numpy.float64)) / cell
    r0 = numpy.floor(r).astype(numpy.int64)
    c0 = numpy.floor(c).astype(numpy.int64)
    fr = (r - r0)[:, None]
    fc = (c - c0)[None, :]
    r1 = numpy.minimum(r0 + 1, coarse.shape[0] - 1)
    c1 = numpy.minimum(c0 + 1, coarse.shape[1] - 1)
    a = coarse[r0][:, c0]
    b = coarse[r0][:, c1]
    cc = coarse[r1][:, c0]
    d = coarse[r1][:, c1]
    return (a * (1 - fr) * (1 - fc) + b * (1 - fr) * fc + cc * fr * (1 - fc) + d * fr * fc)


def synth_v1(nRows, nCols, nBands, seed=0, cell=32, noise=60.0, lo=500.0, hi=4000.0,
        dtype=numpy.uint16, nullFrac=0.0, nullVal=0):
    """
    Smooth field (uniform coarse grid, bilinear upsampled by `cell`) plus Gaussian
    noise, rounded and clipped to [1, max-1] of `dtype`.  With nullFrac > 0 an upper
    left wedge (row+col < sqrt(2*nullFrac*nRows*nCols)) is set to nullVal in all bands.
    Returns a band-sequential (nBands, nRows, nCols) array.
    """
    rng = numpy.random.default_rng(seed)
    info = numpy.iinfo(dtype)
    img = numpy.empty((nBands, nRows, nCols), dtype=dtype)
    for b in range(nBands):
        coarse = rng.uniform(lo, hi, (nRows // cell + 2, nCols // cell + 2))
        fine = _bilinear_zoom(coarse, cell, nRows, nCols)
        fine += rng.normal(0.0, noise, (nRows, nCols))
        img[b] = numpy.clip(numpy.rint(fine), max(info.min, 1), info.max - 1).astype(dtype)
    if nullFrac > 0:
        rr = numpy.arange(nRows)[:, None]
        cc = numpy.arange(nCols)[None, :]
        wedge = (rr + cc) < numpy.sqrt(2.0 * nullFrac * nRows * nCols)
        img[:, wedge] = nullVal
    return img


def synth_tiled(nRows, nCols, nBands, seed=0, cell=32, noise=60.0, lo=500.0, hi=4000.0,
        block=1024, out=None):
    """
    Same statistics as synth_v1 for uint16, built in `block`-sized windows with a
    per-window noise stream, so memory stays at the size of the output.  Any window of
    the result can be regenerated independently (`synth_tiled_window`).
    """
    if out is None:
        out = numpy.empty((nBands, nRows, nCols), dtype=numpy.uint16)
    for r0 in range(0, nRows, block):
        for c0 in range(0, nCols, block):
            h = min(block, nRows - r0)
            w = min(block, nCols - c0)
            out[:, r0:r0 + h, c0:c0 + w] = _synth_block(nRows, nCols, nBands, seed, cell,
                noise, lo, hi, block, r0, c0, h, w)
    return out


def _coarse_grids(nRows, nCols, nBands, seed, cell, lo, hi):
    rng = numpy.random.default_rng(seed)
    return [rng.uniform(lo, hi, (nRows // cell + 2, nCols // cell + 2)) for _ in range(nBands)]


_coarse_cache = {}


def _synth_block(nRows, nCols, nBands, seed, cell, noise, lo, hi, block, r0, c0, h, w):
    key = (nRows, nCols, nBands, seed, cell, lo, hi)
    if key not in _coarse_cache:
        _coarse_cache.clear()
        _coarse_cache[key] = _coarse_grids(nRows, nCols, nBands, seed, cell, lo, hi)
    coarse = _coarse_cache[key]
    rng = numpy.random.default_rng([seed, r0 // block, c0 // block])
    res = numpy.empty((nBands, h, w), dtype=numpy.uint16)
    for b in range(nBands):
        fine = _bilinear_zoom(coarse[b], cell, h, w, r0, c0)
        fine += rng.normal(0.0, noise, (h, w)).astype(numpy.float32)
        res[b] = numpy.clip(numpy.rint(fine), 1, 65534).astype(numpy.uint16)
    return res


def synth_flat(nRows, nCols, nBands, numCells=25, seed=0, border=0, nullVal=65535,
        dtype=numpy.uint16, lo=100, hi=60000):
    """
    Voronoi cells of constant colour (every band constant inside a cell), optional
    null border of `border` pixels.  Cells are far larger than 10001 pixels for the
    sizes used in tests, so this exercises the reference's clump-size cap.
    """
    rng = numpy.random.default_rng(seed)
    cy = rng.uniform(0, nRows, numCells)
    cx = rng.uniform(0, nCols, numCells)
    info = numpy.iinfo(dtype)
    palette = rng.integers(max(lo, info.min + 1), min(hi, info.max - 1), (numCells, nBands))
    rr = numpy.arange(nRows, dtype=numpy.float64)[:, None]
    cc = numpy.arange(nCols, dtype=numpy.float64)[None, :]
    best = numpy.full((nRows, nCols), numpy.inf)
    cellId = numpy.zeros((nRows, nCols), dtype=numpy.int64)
    for i in range(numCells):
        d = (rr - cy[i]) ** 2 + (cc - cx[i]) ** 2
        m = d < best
        best[m] = d[m]
        cellId[m] = i
    img = numpy.empty((nBands, nRows, nCols), dtype=dtype)
    for b in range(nBands):
        img[b] = palette[cellId, b].astype(dtype)
    if border > 0:
        img[:, :border, :] = nullVal
        img[:, -border:, :] = nullVal
        img[:, :, :border] = nullVal
        img[:, :, -border:] = nullVal
    return img


def diagonal_centres(img, numClusters, imgNullVal=None, maxSamples=200000):
    """
    Deterministic cluster centres for benchmarks: a few Lloyd iterations (numpy,
    float64) from centres spread along the diagonal of the data's bounding box, on a
    strided pixel subsample.  The exact centres do not matter for the benchmark, only
    that both arms get the same ones.
    """
    (nBands, nRows, nCols) = img.shape
    x = img.reshape(nBands, -1).T
    step = max(1, x.shape[0] // maxSamples)
    x = x[::step].astype(numpy.float64)
    if imgNullVal is not None:
        x = x[(x != imgNullVal).all(axis=1)]
    mn = x.min(axis=0)
    mx = x.max(axis=0)
    centres = numpy.array([mn + (i + 1) * (mx - mn) / (numClusters + 1)
        for i in range(numClusters)])
    for _ in range(8):
        d = ((x ** 2).sum(axis=1)[:, None] - 2.0 * x @ centres.T +
            (centres ** 2).sum(axis=1)[None, :])
        lab = d.argmin(axis=1)
        for j in range(numClusters):
            m = lab == j
            if m.any():
                centres[j] = x[m].mean(axis=0)
    return centres
