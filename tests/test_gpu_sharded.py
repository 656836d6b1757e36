"""
The sharded tiled path on the GPU: two ranks (gloo process group, both on cuda:0, overlap strips
staged through the host) segment the tiles of one mosaic between them with the CUDA kernels and
stitch them with pyshepseg_b200.distributed.  The windows of both ranks together must be the
mosaic the unmodified reference wrote (tests/golden/tiled_*.npz), with its maxSegId and histogram.
On a multi-GPU box bench.py --gpus N drives the same code with NCCL and device-to-device strips.
"""
import socket

import numpy
import pytest

import goldenutil

pytestmark = pytest.mark.gpu


def _rank_main(rank, world, port, name, resq):
    import torch.distributed as dist
    from pyshepseg_b200 import tiling, distributed, rasterfile, shepseg, timinghooks
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    try:
        c = goldenutil.load(name)
        m = c['meta']
        img = c['img']
        km = goldenutil.Centres(c['centres'])
        (nB, nR, nC) = img.shape
        ti = tiling.getTilesForFile((nC, nR), m['tileSize'], m['overlapSize'])
        msd = shepseg.autoMaxSpectralDiff(km, 'auto', m.get('spectDistPcntile', 50))
        cfg = tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS if rank else tiling.CONC_NONE,
            numWorkers=2 if rank else 0)
        seg = tiling.TiledSegmenter(rasterfile.MemoryRaster(img, nodata=m['imgNullVal']), range(1, nB + 1), ti,
            m['overlapSize'], shepseg._centres(km), m['imgNullVal'], m['fourConnected'], m['minSegmentSize'],
            shepseg.spectralThreshold(msd), m['simpleTileRecode'], cfg, timinghooks.Timers())
        sink = rasterfile.MemorySink(nC, nR)
        comm = distributed.TorchComm()
        (maxSegId, hist) = seg.run(sink, comm)
        total = comm.allreduceSum(sink.array.astype(numpy.int64))     # windows are disjoint
        if rank == 0:
            resq.put((maxSegId, total.astype(numpy.uint32), hist, seg.usedFallback, seg.launches))
    finally:
        dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('name', ['tiled_700x900', 'tiled_null_4x4', 'tiled_640_5x5_8conn', 'tiled_simple_recode'])
def test_two_ranks_equal_reference_mosaic(name):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    resq = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, name, resq)) for r in range(2)]
    for p in procs:
        p.start()
    (maxSegId, mosaic, hist, usedFallback, launches) = resq.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    c = goldenutil.load(name)
    assert launches > 0
    assert int(maxSegId) == c['meta']['maxSegId']
    assert numpy.array_equal(mosaic, c['mosaic'])
    assert numpy.array_equal(numpy.asarray(hist), c['hist'])
    # the offsets came from the per-tile steps the device reported (no replay of the sequential order)
    assert not usedFallback


def _case(name):
    """a golden fixture, or 'row_of_scenes': a seeded raster with two tile rows and many tile
    columns (the shape of bench.py's weak-scaling mosaic), no reference mosaic"""
    if name != 'row_of_scenes':
        return goldenutil.load(name)
    from pyshepseg_b200 import synth
    img = synth.synth_v1(686, 2900, 3, seed=77, cell=14)
    meta = {'tileSize': 256, 'overlapSize': 64, 'minSegmentSize': 20, 'numClusters': 16, 'imgNullVal': None,
        'fourConnected': True, 'simpleTileRecode': False}
    return {'img': img, 'centres': synth.diagonal_centres(img, 16), 'meta': meta}


def _rank_main_api(rank, world, port, name, resq):
    """the same through the public function: every rank hands doTiledShepherdSegmentation the
    WINDOW of the raster its tiles cover (row and column offsets) and a sink of that window"""
    import torch.distributed as dist
    from pyshepseg_b200 import tiling, distributed, rasterfile
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    try:
        c = _case(name)
        m = c['meta']
        img = c['img']
        (nB, nR, nC) = img.shape
        ti = tiling.getTilesForFile((nC, nR), m['tileSize'], m['overlapSize'])
        owner = distributed.partitionTiles(ti, world)
        mine = [(cr, t) for (cr, t) in ti.tiles.items() if owner[cr] == rank]
        (x0, x1) = (min(t[0] for (cr, t) in mine), max(t[0] + t[2] for (cr, t) in mine))
        (y0, y1) = (min(t[1] for (cr, t) in mine), max(t[1] + t[3] for (cr, t) in mine))
        src = rasterfile.MemoryRaster(numpy.ascontiguousarray(img[:, y0:y1, x0:x1]), nodata=m['imgNullVal'],
            yoff=y0, fullYsize=nR, xoff=x0, fullXsize=nC)
        sink = rasterfile.MemorySink(x1 - x0, y1 - y0, yoff=y0, xoff=x0)
        cfg = tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS, numWorkers=1 + rank)
        cfg.comm = distributed.TorchComm()
        res = tiling.doTiledShepherdSegmentation(src, sink, tileSize=m['tileSize'], overlapSize=m['overlapSize'],
            minSegmentSize=m['minSegmentSize'], numClusters=m['numClusters'], imgNullVal=m['imgNullVal'],
            fourConnected=m['fourConnected'], kmeansObj=goldenutil.Centres(c['centres']),
            simpleTileRecode=m['simpleTileRecode'], returnGDALDS=True, concurrencyCfg=cfg)
        full = numpy.zeros((nR, nC), dtype=numpy.int64)
        for ((col, row), (x, y, xs, ys)) in mine:
            (top, bottom, left, right) = tiling.tileMargins(ti, col, row, xs, ys, m['overlapSize'])
            full[y + top:y + bottom, x + left:x + right] = \
                sink.array[y + top - y0:y + bottom - y0, x + left - x0:x + right - x0]
        total = cfg.comm.allreduceSum(full)     # windows are disjoint
        if rank == 0:
            resq.put((int(res.maxSegId), total.astype(numpy.uint32), numpy.asarray(res.outDs.hist),
                res.outDs.metadata.get('STATISTICS_MAXIMUM'), res.gpuLaunches))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('name', ['tiled_700x900', 'tiled_640_5x5_8conn'])
def test_two_ranks_through_the_public_function(name):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    resq = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_rank_main_api, args=(r, 2, port, name, resq)) for r in range(2)]
    for p in procs:
        p.start()
    (maxSegId, mosaic, hist, statMax, launches) = resq.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    c = goldenutil.load(name)
    assert launches > 0
    assert maxSegId == c['meta']['maxSegId']
    assert numpy.array_equal(mosaic, c['mosaic'])
    assert numpy.array_equal(hist, c['hist'])
    assert statMax == repr(c['meta']['maxSegId'])


def test_four_ranks_row_of_scenes_equal_one_rank():
    """four ranks on a mosaic of two tile rows and many columns (a rank owns whole tile columns plus
    a partial one at either end; strips exchanged mid-run) against the same call on one rank"""
    import torch.multiprocessing as mp
    from pyshepseg_b200 import tiling, rasterfile
    c = _case('row_of_scenes')
    m = c['meta']
    ti = tiling.getTilesForFile((c['img'].shape[2], c['img'].shape[1]), m['tileSize'], m['overlapSize'])
    assert ti.nrows == 2 and ti.ncols >= 12
    one = tiling.doTiledShepherdSegmentation(rasterfile.MemoryRaster(c['img']), None, tileSize=m['tileSize'],
        overlapSize=m['overlapSize'], minSegmentSize=m['minSegmentSize'], numClusters=m['numClusters'],
        fourConnected=True, kmeansObj=goldenutil.Centres(c['centres']), outputDriver='MEM', returnGDALDS=True)
    ctx = mp.get_context('spawn')
    resq = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_rank_main_api, args=(r, 4, port, 'row_of_scenes', resq)) for r in range(4)]
    for p in procs:
        p.start()
    (maxSegId, mosaic, hist, statMax, launches) = resq.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert maxSegId == int(one.maxSegId) and maxSegId > 1000
    assert numpy.array_equal(mosaic, one.outDs.array)
    assert numpy.array_equal(hist, numpy.asarray(one.outDs.hist))

