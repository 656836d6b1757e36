"""
The sharded stitch (pyshepseg_b200.distributed) on the CPU: tiles of one mosaic dealt over 2 and 3
ranks of a gloo process group, the device work replaced by numpy (the same per-tile tables the
GPU kernels compute, emulated in test_stitch_host.numpy_tile_tables).  Every rank's windows put
together must be the mosaic the oracle's restatement of tiling.stitchTiles builds sequentially,
with the same maxSegId -- in the fast path (offsets from an all-gather of counts, lazy look-ups
into gathered neighbour tables), in the forced sequential fall-back and with simpleTileRecode.
"""
import os
import socket

import numpy
import pytest

from oracle import oracle
from pyshepseg_b200 import tiling, distributed
from test_stitch_host import numpy_tile_tables


def blobby_tiles(seed, nR, nC, tileSize, overlap, cell=6, nLabels=40):
    rng = numpy.random.default_rng(seed)
    ti = tiling.getTilesForFile((nC, nR), tileSize, overlap)
    segs = {}
    for cr in distributed.rowMajor(ti):
        (x, y, xs, ys) = ti.tiles[cr]
        coarse = rng.integers(0, nLabels, (ys // cell + 2, xs // cell + 2))
        lab = numpy.kron(coarse, numpy.ones((cell, cell), dtype=numpy.int64))[:ys, :xs]
        (_, inv) = numpy.unique(lab, return_inverse=True)
        segs[cr] = inv.reshape(ys, xs).astype(numpy.uint32)     # contiguous ids, 0 = null
    return (ti, segs)


def segmented_tiles(seed, nR, nC, tileSize, overlap, minSeg=12, k=14):
    """labels as the path produces them: the oracle's segmentation of every tile of a synthetic image"""
    from pyshepseg_b200 import synth
    import goldenutil
    img = synth.synth_v1(nR, nC, 3, seed=seed, cell=16)
    km = goldenutil.Centres(synth.diagonal_centres(img, k))
    ti = tiling.getTilesForFile((nC, nR), tileSize, overlap)
    segs = {}
    for cr in distributed.rowMajor(ti):
        (x, y, xs, ys) = ti.tiles[cr]
        sub = numpy.ascontiguousarray(img[:, y:y + ys, x:x + xs])
        segs[cr] = oracle.doShepherdSegmentation(sub, minSegmentSize=minSeg, kmeansObj=km).segimg
    return (ti, segs)


def make_tiles(kind, seed, nR, nC, tileSize, overlap):
    return (blobby_tiles if kind == 'blobby' else segmented_tiles)(seed, nR, nC, tileSize, overlap)


class NumpyOps(object):
    """what tiling.TiledSegmenter does on the GPU, in numpy"""
    def __init__(self, ti, segs, overlap, out):
        import torch
        self.torch = torch
        (self.ti, self.segs, self.ov, self.out) = (ti, segs, overlap, out)

    def sendStrip(self, cr, which):
        t = self.segs[cr]
        s = t[-self.ov:, :] if which == 'bottom' else t[:, -self.ov:]
        return self.torch.from_numpy(numpy.ascontiguousarray(s).astype(numpy.int32))

    def recvStrip(self, cr, which, shape):
        return self.torch.zeros(shape, dtype=self.torch.int32)

    def _strip(self, cr, nb, which, given):
        if given is None:
            return None
        if isinstance(given, str):
            t = self.segs[nb]
            return t[-self.ov:, :] if which == 'bottom' else t[:, -self.ov:]
        return given.numpy().astype(numpy.uint32)

    def tables(self, cr, top, left):
        (x, y, xs, ys) = self.ti.tiles[cr]
        m = tiling.tileMargins(self.ti, cr[0], cr[1], xs, ys, self.ov)
        topB = self._strip(cr, (cr[0], cr[1] - 1), 'bottom', top)
        leftB = self._strip(cr, (cr[0] - 1, cr[1]), 'right', left)
        (tb, rank, flags, pk, pc) = numpy_tile_tables(self.segs[cr], self.ov, topB, leftB, *m)
        return distributed.TileTable(tb.maxId, tb.countNew, rank, flags, pk, pc)

    def apply(self, cr, lut, tb):
        (x, y, xs, ys) = self.ti.tiles[cr]
        (top, bottom, left, right) = tiling.tileMargins(self.ti, cr[0], cr[1], xs, ys, self.ov)
        self.out[y + top:y + bottom, x + left:x + right] = lut[self.segs[cr][top:bottom, left:right]]


class NumpyRelOps(NumpyOps):
    """the same with the lut put together from rank + offset and the crossing segments' ids, as
    ssg_apply_rel_lut_device does"""
    def applyRel(self, cr, offset, crossLabels, crossIds, tb):
        numbered = (tb.flags & 2) != 0
        lut = numpy.where(numbered, tb.rank + numpy.uint32(offset), 0).astype(numpy.uint32)
        if self.simple:
            lut = numpy.arange(tb.maxId + 1, dtype=numpy.uint32) + numpy.uint32(offset)
            lut[0] = 0
        lut[crossLabels] = crossIds
        self.apply(cr, lut, tb)


def _rank_main(rank, world, port, case, forceSequential, resq):
    import torch.distributed as dist
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    try:
        (kind, seed, nR, nC, tileSize, overlap, simple) = case
        (ti, allSegs) = make_tiles(kind, seed, nR, nC, tileSize, overlap)
        comm = distributed.TorchComm()
        if world >= 4:
            # force the sizes-first gather: the request arrays do not fit the one-shot buffer
            realGather = comm.allgatherArray
            comm.allgatherArray = lambda a, quick=64: realGather(a, quick)
        st = distributed.ShardedStitch(ti, overlap, simple, comm)
        st.forceSequential = forceSequential
        # a rank only ever touches the labels of its own tiles (remote strips arrive by message)
        segs = dict((cr, allSegs[cr]) for cr in st.mine)
        out = numpy.zeros((nR, nC), dtype=numpy.uint32)
        ops = NumpyOps(ti, segs, overlap, out) if seed % 2 else NumpyRelOps(ti, segs, overlap, out)
        ops.simple = simple
        (maxSegId, offsets, luts) = st.run(ops)
        total = comm.allreduceSum(out.astype(numpy.int64))    # windows are disjoint
        if rank == 0:
            resq.put((maxSegId, total.astype(numpy.uint32), st.usedFallback, len(st.mine)))
    finally:
        dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


CASES = [
    # name, world, (labels, seed, nR, nC, tileSize, overlap, simple), fallback expected
    ('w2_4x3', 2, ('segmented', 11, 260, 300, 96, 32, False), False),
    ('w3_5x5', 3, ('segmented', 12, 330, 340, 96, 40, False), False),
    ('w2_simple', 2, ('segmented', 13, 260, 300, 96, 32, True), False),
    # labels that are not connected segments (every blob of one value is one "segment"): ids
    # numbered without a pixel in the trimmed window, votes for 0, ties
    ('w2_blobby', 2, ('blobby', 14, 260, 300, 96, 32, False), False),
    # the sequential fall-back (taken when a window holds an inherited id above the running
    # maximum), forced
    ('w2_forced_sequential', 2, ('blobby', 14, 260, 300, 96, 32, False), True),
    ('w3_blobby_simple', 3, ('blobby', 15, 260, 300, 96, 32, True), False),
    # more ranks than tile rows: a rank's upper neighbours belong to two different ranks, ids are
    # inherited around corners across rank boundaries (several look-up rounds)
    ('w4_6x6', 4, ('segmented', 16, 420, 430, 96, 32, False), False),
    ('w8_blobby_7x6', 8, ('blobby', 17, 400, 460, 96, 32, False), False),
    # a row of scenes (bench.py's weak-scaling mosaic): two tile rows, many tile columns, a rank owns
    # whole columns plus a partial one at either end
    ('w8_row_of_scenes', 8, ('segmented', 18, 343, 1900, 128, 32, False), False),
    ('w4_row_of_scenes_blobby', 4, ('blobby', 19, 343, 1300, 128, 32, False), False),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_sharded_stitch_gloo(case):
    import torch.multiprocessing as mp
    (name, world, prm, fallbackExpected) = case
    (kind, seed, nR, nC, tileSize, overlap, simple) = prm
    ctx = mp.get_context('spawn')
    resq = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, prm, fallbackExpected, resq))
        for r in range(world)]
    for p in procs:
        p.start()
    (maxSegId, mosaic, usedFallback, nMine) = resq.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (ti, segs) = make_tiles(kind, seed, nR, nC, tileSize, overlap)
    assert 0 < nMine < ti.getNumTiles()
    (want, wantMax, _) = oracle.stitchTiles(segs, ti, nC, nR, overlap, simpleTileRecode=simple)
    assert maxSegId == wantMax
    assert numpy.array_equal(mosaic, want)
    assert usedFallback == fallbackExpected


def _costs(ti, owner, world):
    cost = numpy.zeros(world)      # the partition balances the modelled segmentation cost
    for (cr, t) in ti.tiles.items():
        cost[owner[cr]] += distributed.TILE_COST_FIXED + distributed.TILE_COST_PER_MPIX * t[2] * t[3] / 1e6
    return cost


def test_partition_is_blocks_and_balanced():
    ti = tiling.getTilesForFile((40000, 40000), 4096, 1024)
    for world in (1, 2, 4, 8):
        owner = distributed.partitionTiles(ti, world)
        assert set(owner.values()) == set(range(world))
        for rank in range(world):
            # a band of tile rows, and in it whole tile columns plus at most a partial column at
            # either end: contiguous in the band's column-by-column order
            mine = [cr for cr in ti.tiles if owner[cr] == rank]
            rows = sorted(set(r for (c, r) in mine))
            assert rows == list(range(rows[0], rows[-1] + 1))
            band = [(c, r) for c in range(ti.ncols) for r in rows]
            pos = sorted(band.index(cr) for cr in mine)
            assert pos == list(range(pos[0], pos[-1] + 1))
        cost = _costs(ti, owner, world)
        assert cost.max() / cost.mean() < 1.04
    # no grid of 7 pieces fits 3 x 3 tiles: contiguous chunks of the row-major list
    ti = tiling.getTilesForFile((12000, 12000), 4096, 512)
    assert (ti.nrows, ti.ncols) == (3, 3)
    owner = distributed.partitionTiles(ti, 7)
    assert set(owner.values()) == set(range(7))


def test_chunk_partition_is_contiguous_and_balanced():
    ti = tiling.getTilesForFile((40000, 40000), 4096, 1024)
    order = distributed.rowMajor(ti)
    for world in (1, 2, 4, 8):
        owner = distributed.partitionChunks(ti, world)
        ranks = [owner[cr] for cr in order]
        assert ranks == sorted(ranks) and set(ranks) == set(range(world))
        cost = _costs(ti, owner, world)
        assert cost.max() / cost.mean() < 1.05


def test_lazy_resolver_equals_sequential_single_rank():
    """world of one: LocalComm, every look-up local"""
    (nR, nC, tileSize, overlap) = (300, 280, 96, 32)
    (ti, segs) = segmented_tiles(21, nR, nC, tileSize, overlap)
    out = numpy.zeros((nR, nC), dtype=numpy.uint32)
    st = distributed.ShardedStitch(ti, overlap, False, distributed.LocalComm())
    (maxSegId, offsets, luts) = st.run(NumpyOps(ti, segs, overlap, out))
    (want, wantMax, _) = oracle.stitchTiles(segs, ti, nC, nR, overlap)
    assert maxSegId == wantMax and numpy.array_equal(out, want)
    assert not st.usedFallback


def test_table_pack_roundtrip():
    (ti, segs) = blobby_tiles(5, 200, 220, 96, 32)
    ops = NumpyOps(ti, segs, 32, None)
    tabs = {}
    for cr in distributed.rowMajor(ti):
        top = 'local' if cr[1] > 0 else None
        left = 'local' if cr[0] > 0 else None
        tabs[cr] = ops.tables(cr, top, left)
    back = distributed.unpackTables(distributed.packTables(tabs))
    assert sorted(back) == sorted(tabs)
    for cr in tabs:
        for f in ('rank', 'flags', 'pairKeys', 'pairCounts'):
            assert getattr(back[cr], f).dtype == getattr(tabs[cr], f).dtype
            assert numpy.array_equal(getattr(back[cr], f), getattr(tabs[cr], f))
        assert (back[cr].maxId, back[cr].countNew) == (tabs[cr].maxId, tabs[cr].countNew)
