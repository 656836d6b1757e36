"""
bench.py -- Mpixels/s of the full Shepherd segmentation (assign + clump + eliminate + stitch)
on B200, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c4_40000]

A step is one pass of the hot path over one synthetic raster: the workload named by
BASELINE.json configs[1], a Sentinel-2-like 10980 x 10980 x 4 uint16 raster segmented by
doTiledShepherdSegmentation with tileSize=4096, overlapSize=1024 (2 x 2 tiles of 4096 and 7908
pixels), numClusters=60, minSegmentSize=50, maxSpectralDiff='auto', four-connected, with given
cluster centres (the k-means fit is host-side set-up, not part of the path).  Pixels counted are
the unique pixels of the raster.

Two numbers per run:
  value  the tiled segmentation with the raster resident in HBM and the mosaic written to HBM
         (pyshepseg_b200.tiling.TiledSegmenter over DeviceRaster / DeviceMosaicSink);
  e2e    the same call with the raster in pinned HOST memory and the mosaic delivered to HOST
         memory, copies inside the timed region (the public path of doTiledShepherdSegmentation
         with two segmentation workers overlapping copies and kernels).

With --gpus N (launched by torchrun, one rank per GPU) the raster is ONE mosaic of N such scenes
(weak scaling), its tiles dealt over the ranks (pyshepseg_b200.distributed): every rank holds the
window of the raster its tiles cover, segments them, and the ranks stitch the mosaic with global
ids -- overlap strips device to device over NCCL, two or three small all-gathers, no data-path
collective; NCCL also broadcasts the cluster centres from rank 0.  After the timed regions the
N-rank mosaic is compared with rank 0 segmenting the same raster alone (parity_vs_one_rank).
--workload c4_40000 is BASELINE config 4 instead: one 40000 x 40000 x 4 mosaic whatever N
(strong scaling, 144 tiles).

Also in the line: roofline (dominant kernel by event-timed share, algorithmic bytes against the
measured HBM peak, ncu DRAM traffic), roofline_other / roofline_stages, kernels (event pair per
launch on one stream), cpu_baseline (numba reference) and cpu_baseline_port, clocks, and
parity_vs_oracle: the e2e mosaic of the last timed step against the CPU oracle on the same raster.

--impl reference times the reference itself on the host cores: the UNMODIFIED pyshepseg 2.0.3
(installed from /root/reference into the git-ignored baseline/_ref/, numba + scikit-learn) runs
shepseg.doShepherdSegmentation in one process per core, each on its own crop of the workload
raster, JIT warmed beforehand (cpu_baseline.kind "reference").  Where that install is missing it
falls back to the C port of the path in oracle/ (kind "port"), which the GPU arm also times
beside the numba figure.  Both are bounded samples of the same workload; pixels are credited as
unique mosaic pixels (tile pixels divided by the workload's tile-pixel / unique-pixel ratio).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pyshepseg_b200 import synth  # noqa: E402

WORKLOAD = {
    'name': 's2like_10980x10980x4_u16_tiled',
    'rows': 10980, 'cols': 10980, 'bands': 4, 'tileSize': 4096, 'overlapSize': 1024,
    'numClusters': 60, 'minSegmentSize': 50, 'maxSpectralDiff': 'auto', 'fourConnected': True,
}
METRIC = 'Mpixels/sec full segmentation (assign+clump+eliminate+stitch), tiled 10980x10980x4 uint16'
UNIT = 'Mpixel/s'
E2E_WORKERS = int(os.environ.get('BENCH_E2E_WORKERS', '3'))   # segmentation workers of the host-to-host run
# scenes (rows, cols) of the weak-scaling mosaic per N: side by side.  A row of scenes keeps two tile rows
# (4096 and 7908 pixels high, as in the single scene), 1.40-1.43 tile pixels per unique pixel and two
# rim tiles per rank boundary; the squarer 2 x 2 / 2 x 4 layouts measured in profiles/r2_scaling.md have
# six tile rows, 1.52-1.58 tile pixels per unique pixel and six rim tiles per boundary
SCENE_GRID = {1: (1, 1), 2: (1, 2), 4: (1, 4), 8: (1, 8)}
RESIDENT_WORKERS = int(os.environ.get('BENCH_RESIDENT_WORKERS', '2'))   # 0: segment and stitch in one thread

# algorithmic bytes per pixel of the kernels:
# what one launch must move at the very least, per pixel it processes
def algorithmic_bytes_per_pixel(kernel, nB):
    table = {
        'k_assign': 2 * nB + 4,            # read the bands, write the int32 cluster
        'k_assign_grid': 2 * nB + 4,
        'k_pixel_pass': 2 * nB + 8,        # bands once, label read (pending relabel folded in) + written
        'k_apply_lut_extents': 8,          # label read + final label written (extent tables are per segment)
        'k_ccl_flatten_count': 8,
        # the region-merge kernels only touch segment records (32 B each), a few bytes per pixel;
        # they are charged the whole "eliminate small segments" figure of SURVEY 8(d) -- bands
        # once, labels read + written -- because that stage is what they dominate
        'k_merge_wide': 2 * nB + 8,
        'k_merge_lean': 2 * nB + 8,
        'k_merge_tail': 2 * nB + 8,
        'k_ccl_local': 8,                  # read cluster, write root label
        'k_ccl_flatten': 8,
        'k_gather_ids': 8,
        'k_single_decide': 8,              # read label + size of every pixel
        'k_band_sums': 2 * nB + 4,         # read the bands and the label
        'k_list_fill': 4,
        'k_small_persistent': 2 * nB + 8,  # the "eliminate" figure: bands once, labels read + written
        'k_apply_lut': 8,
        'k_apply_lut_window': 8,
        'k_tile_extents': 4,
        'k_tile_max': 4,
        'k_count_roots': 4,
        'k_number_roots': 4,
    }
    return table.get(kernel)


# DRAM bytes per pixel (dram__bytes_read.sum + dram__bytes_write.sum of `ncu --set full` captures on
# a 4096 x 4096 x 4 tile, profiles/r2_tile4096_ncu_full.md, divided by the tile's pixels): scaled by
# the pixels one launch processes it gives `roofline.traffic`
NCU_DRAM_BYTES_PER_PIXEL = {
    'k_merge_wide': (96.4e6 + 20.1e6) / 4096 ** 2,
    'k_merge_lean': (52.9e6 + 2.1e6) / 4096 ** 2,
    'k_assign_grid': (135.4e6 + 41.5e6) / 4096 ** 2,
    'k_ccl_local': (67.2e6 + 26.8e6) / 4096 ** 2,
    'k_pixel_pass': (219.6e6 + 63.2e6) / 4096 ** 2,
    'k_gather_ids': (122.7e6 + 35.3e6) / 4096 ** 2,
    'k_ccl_flatten_count': (67.1e6 + 6.5e6) / 4096 ** 2,
}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return (json.load(open(path)).get('hbm_gbs', 6650.0), 'measured')
    return (6650.0, 'fallback')


class KM(object):
    def __init__(self, centres):
        self.cluster_centers_ = numpy.ascontiguousarray(centres, dtype=numpy.float64)


def make_scene(wl, seed, out=None):
    """The synthetic raster of one scene (cached on /tmp so the two arms share it)."""
    cache = '/tmp/shepseg_bench_%s_seed%d.npy' % (wl['name'], seed)
    shape = (wl['bands'], wl['rows'], wl['cols'])
    if os.path.exists(cache):
        try:
            a = numpy.load(cache, mmap_mode='r')
            if a.shape == shape:
                if out is None:
                    return numpy.array(a)
                out[...] = a
                return out
        except Exception:
            pass
    img = synth.synth_tiled(wl['rows'], wl['cols'], wl['bands'], seed=seed, out=out)
    try:
        numpy.save(cache + '.tmp.npy', img)
        os.replace(cache + '.tmp.npy', cache)
    except Exception:
        pass
    return img


def scene_centres(wl, img, seed=1):
    """
    Cluster centres of the workload: scikit-learn KMeans fitted the way the reference fits it
    (fitSpectralClusters with fixedKMeansInit=True, shepseg.py:252-314) on a 1 % subsample of the
    raster (every 10th row and column), cached on /tmp so that both arms use the same numbers.
    """
    cache = '/tmp/shepseg_bench_%s_seed%d_centres_k%d_%dx%d.npy' % (wl['name'], seed, wl['numClusters'],
        img.shape[1], img.shape[2])
    if os.path.exists(cache):
        try:
            c = numpy.load(cache)
            if c.shape == (wl['numClusters'], img.shape[0]):
                return c
        except Exception:
            pass
    from pyshepseg_b200 import shepseg
    sub = numpy.ascontiguousarray(img[:, ::10, ::10])
    km = shepseg.fitSpectralClusters(sub, wl['numClusters'], 100, None, True)
    c = numpy.ascontiguousarray(km.cluster_centers_, dtype=numpy.float64)
    try:
        numpy.save(cache + '.tmp.npy', c)
        os.replace(cache + '.tmp.npy', cache)
    except Exception:
        pass
    return c


class ClockSampler(object):
    """SM clock, power and throttle reasons during the timed region (B200_PROFILING.md), read
    through NVML in a thread of this process every 50 ms (an nvidia-smi child polling the driver
    was seen to stall CUDA calls of the benchmark for tens of milliseconds)."""
    def __init__(self, device):
        self.device = device
        self.samples = []
        self.stopFlag = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            # the first call of every query is slow (tens of ms) and holds up CUDA calls of this
            # process meanwhile: make them here, before the timed region starts
            n = pynvml
            n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            (n.nvmlDeviceGetCurrentClocksEventReasons if hasattr(n, 'nvmlDeviceGetCurrentClocksEventReasons')
                else n.nvmlDeviceGetCurrentClocksThrottleReasons)(self.handle)
            n.nvmlDeviceGetPowerUsage(self.handle)
            n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _run(self):
        n = self.nvml
        while not self.stopFlag.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n,
                    'nvmlDeviceGetCurrentClocksEventReasons') else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((sm, reasons, power))
            except Exception:
                pass
            self.stopFlag.wait(0.05)

    def stop(self):
        if self.nvml is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml unavailable']}
        self.stopFlag.set()
        self.thread.join(timeout=2)
        n = self.nvml
        try:
            smMax = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:
            smMax = None
        names = (('hw_slowdown', 0x8), ('sw_thermal_slowdown', 0x20), ('hw_thermal_slowdown', 0x40),
            ('hw_power_brake_slowdown', 0x80), ('sw_power_cap', 0x4))
        reasons = set()
        for (_, r, _) in self.samples:
            for (name, bit) in names:
                if r & bit:
                    reasons.add(name)
        sm = [s[0] for s in self.samples]
        power = [s[2] for s in self.samples]
        return {'sm_mhz': float(numpy.median(sm)) if sm else None, 'sm_max_mhz': smMax,
            'reasons': sorted(reasons), 'samples': len(sm),
            'power_w_max': max(power) if power else None}


# ---------------------------------------------------------------------------------------------
# the CPU arm: the unmodified reference (numba) on the host cores; the C port beside it
# ---------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def tile_pixel_ratio(tileInfo, nR, nC):
    """tile pixels per unique pixel of the workload's tile layout (1.195 for 10980 / 4096 / 1024)"""
    return sum(t[2] * t[3] for t in tileInfo.tiles.values()) / float(nR * nC)


def crop_origins(wl, n, tile):
    out = []
    step = max(1, (wl['rows'] - tile) // max(1, n - 1)) if n > 1 else 0
    for i in range(n):
        y = min(i * step, wl['rows'] - tile)
        x = min((i * 977) % max(1, wl['cols'] - tile), wl['cols'] - tile)
        out.append((y, x))
    return out


def cpu_port_sample(wl, img, centres, threads, tile=4096):
    """
    One bounded sample of the workload through the C port of the path (oracle/): `threads` tiles
    of tile x tile pixels cut from the raster, one per thread (the reference's numba stages are
    single threaded; using all cores means one tile per core).
    Returns (tile pixels, seconds).
    """
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import oracle
    oracle.lib()
    km = KM(centres)
    crops = [numpy.ascontiguousarray(img[:, y:y + tile, x:x + tile]) for (y, x) in crop_origins(wl, threads, tile)]
    done = [None] * threads

    def work(i):
        done[i] = oracle.doShepherdSegmentation(crops[i], minSegmentSize=wl['minSegmentSize'],
            maxSpectralDiff=wl['maxSpectralDiff'], fourConnected=wl['fourConnected'], kmeansObj=km)
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    t0 = time.time()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.time() - t0
    return (threads * tile * tile, dt)


def _numba_worker(conn, refDir, scenePath, origin, tile, wlDict, centres):
    """One process of the numba arm: imports the unmodified reference from baseline/_ref, warms
    the JIT on a 64 x 64 image, then segments its own crop every time it is told to."""
    try:
        os.environ['OMP_NUM_THREADS'] = '1'          # predict's OpenMP team: one core per process
        os.environ['NUMBA_NUM_THREADS'] = '1'
        sys.path.insert(0, refDir)
        from pyshepseg import shepseg as ref
        from sklearn.cluster import KMeans
        import warnings
        warnings.filterwarnings('ignore')
        k = centres.shape[0]
        km = KMeans(n_clusters=k, n_init=1, init=centres, max_iter=1)
        km.fit(centres)                              # a fitted object whose predict() works ...
        km.cluster_centers_ = numpy.ascontiguousarray(centres, dtype=numpy.float64)   # ... with OUR centres
        scene = numpy.load(scenePath, mmap_mode='r')
        (y, x) = origin
        crop = numpy.ascontiguousarray(scene[:, y:y + tile, x:x + tile])
        ref.doShepherdSegmentation(numpy.ascontiguousarray(crop[:, :64, :64]), minSegmentSize=wlDict['minSegmentSize'],
            maxSpectralDiff=wlDict['maxSpectralDiff'], fourConnected=wlDict['fourConnected'], kmeansObj=km)
        conn.send(('ready', ref.__file__))
        while True:
            msg = conn.recv()
            if msg == 'stop':
                break
            size = int(msg)
            sub = crop if size >= tile else numpy.ascontiguousarray(crop[:, :size, :size])
            t0 = time.time()
            res = ref.doShepherdSegmentation(sub, minSegmentSize=wlDict['minSegmentSize'],
                maxSpectralDiff=wlDict['maxSpectralDiff'], fourConnected=wlDict['fourConnected'], kmeansObj=km)
            conn.send(('done', time.time() - t0, int(res.segimg.max())))
    except Exception as e:      # noqa
        import traceback
        conn.send(('error', traceback.format_exc()))


class NumbaPool(object):
    """`procs` processes, each holding the reference and one crop of the scene."""
    def __init__(self, wl, scenePath, centres, procs, tile):
        import multiprocessing
        mp = multiprocessing.get_context('spawn')
        self.procs = []
        self.tile = tile
        wlDict = dict((k, wl[k]) for k in ('minSegmentSize', 'maxSpectralDiff', 'fourConnected'))
        for origin in crop_origins(wl, procs, tile):
            (a, b) = mp.Pipe()
            pr = mp.Process(target=_numba_worker, args=(b, REF_DIR, scenePath, origin, tile, wlDict, centres),
                daemon=True)
            pr.start()
            self.procs.append((pr, a))
        self.refFile = None
        for (pr, a) in self.procs:
            msg = a.recv()
            if msg[0] != 'ready':
                self.close()
                raise RuntimeError('numba worker failed: %s' % (msg[1],))
            self.refFile = msg[1]

    def step(self, size=None):
        """every process segments its crop (or its top-left size x size corner); (tile pixels, seconds)"""
        size = self.tile if size is None else min(size, self.tile)
        t0 = time.time()
        for (pr, a) in self.procs:
            a.send(size)
        for (pr, a) in self.procs:
            msg = a.recv()
            if msg[0] != 'done':
                self.close()
                raise RuntimeError('numba worker failed: %s' % (msg[1],))
        return (len(self.procs) * size * size, time.time() - t0)

    def close(self):
        for (pr, a) in self.procs:
            try:
                a.send('stop')
            except Exception:
                pass
        for (pr, a) in self.procs:
            pr.join(timeout=5)
            if pr.is_alive():
                pr.kill()
        self.procs = []


def scene_path(wl, seed=1):
    return '/tmp/shepseg_bench_%s_seed%d.npy' % (wl['name'], seed)


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = min(cores, 64)      # one tile per host core (the reference's numba stages are single threaded)
    img = make_scene(wl, seed=1)
    centres = scene_centres(wl, img)
    from pyshepseg_b200 import tiling
    (gr, gc) = (1, 1) if wl.get('strong') else SCENE_GRID.get(args.gpus, (1, args.gpus))
    tileInfo = tiling.getTilesForFile((wl['cols'] * gc, wl['rows'] * gr), wl['tileSize'], wl['overlapSize'])
    ratio = tile_pixel_ratio(tileInfo, wl['rows'] * gr, wl['cols'] * gc)
    haveRef = os.path.isdir(os.path.join(REF_DIR, 'pyshepseg')) and os.path.exists(scene_path(wl)) \
        and not args.port_only
    pix = 0
    secs = 0.0
    if haveRef:
        # a step = one crop per core through the unmodified reference; crops of 2048 pixels keep a
        # driver-length run (25 steps) within a few minutes at ~3 Mpixel/s per core
        tile = 1024 if args.quick else 2048
        pool = NumbaPool(wl, scene_path(wl), centres, procs, tile)
        try:
            for _ in range(args.warmup):
                pool.step(256)
            for _ in range(args.steps):
                (p, s) = pool.step()
                pix += p
                secs += s
        finally:
            pool.close()
        kind = 'reference'
        sample = ('%d crops of %dx%dx%d of the workload raster per step, one per process, through the unmodified '
            'pyshepseg %s doShepherdSegmentation (numba JIT warmed; no stitch); pixels credited as unique mosaic '
            'pixels = tile pixels / %.3f' % (procs, tile, tile, wl['bands'], '2.0.3', ratio))
    else:
        tile = 2048 if args.quick else 4096
        for _ in range(args.warmup):
            cpu_port_sample(wl, img, centres, procs, tile=min(tile, 1024))
        for _ in range(args.steps):
            (p, s) = cpu_port_sample(wl, img, centres, procs, tile=tile)
            pix += p
            secs += s
        kind = 'port'
        sample = ('%d tiles of %dx%dx%d cut from the workload raster per step, one per thread, through the C port '
            'of the path (oracle/; baseline/_ref not installed); pixels credited as unique mosaic pixels = tile '
            'pixels / %.3f' % (procs, tile, tile, wl['bands'], ratio))
    value = pix / ratio / secs / 1e6
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': secs / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u16',
        'data': 'synthetic', 'config': workload_config(wl, args.gpus, None, tileInfo),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, gpus, shape=None, tileInfo=None):
    (gr, gc) = (1, 1) if wl.get('strong') else SCENE_GRID.get(gpus, (1, gpus))
    (nR, nC) = shape if shape is not None else (wl['rows'] * gr, wl['cols'] * gc)
    name = wl['name'] if (gpus == 1 or wl.get('strong')) else '%s_mosaic_of_%dx%d_scenes' % (wl['name'], gr, gc)
    tiles = ('%dx%d' % (tileInfo.nrows, tileInfo.ncols)) if tileInfo is not None else None
    return {'workload': name, 'raster': [wl['bands'], nR, nC], 'dtype': 'uint16',
        'tileSize': wl['tileSize'], 'overlapSize': wl['overlapSize'], 'tiles': tiles,
        'numClusters': wl['numClusters'], 'minSegmentSize': wl['minSegmentSize'],
        'maxSpectralDiff': wl['maxSpectralDiff'], 'fourConnected': wl['fourConnected'],
        'scenes': 1 if wl.get('strong') else gpus,
        'parallelism': (('one mosaic, the same for every N, ' if wl.get('strong') else 'one mosaic of %d scenes, ' % gpus) +
            'tiles dealt over %d GPUs in blocks of tile columns of equal cost, overlap strips over NCCL, ids global' % gpus)
            if gpus > 1 else 'single GPU',
        'l2_policy': 'inputs larger than L2 (964 MB raster, 482 MB mosaic per scene)'}


# ---------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    from pyshepseg_b200 import _lib, shepseg, tiling, rasterfile, timinghooks, distributed

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('--gpus %d needs torchrun with %d ranks' % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a rank that dies must take the job down quickly instead of leaving the others in a
        # collective for ten minutes
        dist.init_process_group('nccl', device_id=torch.device('cuda', local),
            timeout=datetime.timedelta(seconds=240))

    # ---- set-up (untimed): the raster in pinned host memory, centres from rank 0 over NCCL ----
    # N = 1: the scene.  N > 1: ONE mosaic of N scenes (weak scaling), its tiles dealt over the
    # ranks in contiguous row-major chunks; a rank holds the row band its tiles cover.
    nB = wl['bands']
    (gr, gc) = (1, 1) if wl.get('strong') else SCENE_GRID.get(world, (1, world))
    (nR, nC) = (wl['rows'] * gr, wl['cols'] * gc)
    tileInfo = tiling.getTilesForFile((nC, nR), wl['tileSize'], wl['overlapSize'])
    comm = None
    (y0, y1) = (0, nR)
    (x0, x1) = (0, nC)
    if world > 1:
        comm = distributed.TorchComm(torch.device('cuda', local))
        owner = distributed.partitionTiles(tileInfo, world)
        mineT = [tileInfo.tiles[cr] for cr in tileInfo.tiles if owner[cr] == rank]
        (y0, y1) = (min(t[1] for t in mineT), max(t[1] + t[3] for t in mineT))
        (x0, x1) = (min(t[0] for t in mineT), max(t[0] + t[2] for t in mineT))
    (bandRows, bandCols) = (y1 - y0, x1 - x0)      # the window of the raster this rank's tiles cover
    pinnedImg = _lib.PinnedArray((nB, bandRows, bandCols), numpy.uint16)
    if wl.get('strong'):
        # the 40000 x 40000 mosaic: one synthetic 10000 x 10000 patch (rank 0 makes it, NCCL hands
        # it round) laid out 4 x 4 with every other copy mirrored, so that the raster is continuous
        # across the seams; a rank cuts its own window out of the patch on its GPU
        P = min(10000, nR)
        patch = torch.empty((nB, P, P), dtype=torch.int16, device='cuda')
        if rank == 0:
            patch.copy_(torch.from_numpy(synth.synth_tiled(P, P, nB, seed=1).view(numpy.int16)))
        if dist is not None:
            dist.broadcast(patch.view(torch.uint8), 0)

        def mirrored(lo, hi):
            i = torch.arange(lo, hi, device='cuda') % (2 * P)
            return torch.where(i < P, i, 2 * P - 1 - i)
        (ri, ci) = (mirrored(y0, y1), mirrored(x0, x1))
        img = pinnedImg.array
        for b in range(nB):
            img[b] = patch[b].index_select(0, ri).index_select(1, ci).cpu().numpy().view(numpy.uint16)
        del patch
        torch.cuda.empty_cache()
    elif world == 1:
        img = make_scene(wl, seed=1, out=pinnedImg.array)
    else:
        img = synth.synth_window(nR, nC, nB, 1, y0, x0, bandRows, bandCols, out=pinnedImg.array)
    if rank == 0:
        centres = scene_centres(wl, img)
    else:
        centres = numpy.zeros((wl['numClusters'], nB))
    if dist is not None:
        t = torch.from_numpy(centres).cuda()
        dist.broadcast(t, 0)
        centres = t.cpu().numpy()
    centres = numpy.ascontiguousarray(centres, dtype=numpy.float64)
    km = KM(centres)
    msd = shepseg.autoMaxSpectralDiff(km, wl['maxSpectralDiff'], 50)
    thr = shepseg.spectralThreshold(msd)
    pinnedOut = _lib.PinnedArray((bandRows, bandCols), numpy.uint32)

    state = tiling.gpuState(local)
    ctx0 = state.slot(0).ctx
    devImg = ctx0.dev_alloc(img.nbytes)
    ctx0.call('ssg_memcpy_h2d', devImg, _lib.ptr(img), img.nbytes)
    ctx0.synchronize()
    devMosaic = ctx0.dev_alloc(bandRows * bandCols * 4)

    def step_resident(profile=False):
        cfg = tiling.SegmentationConcurrencyConfig(devices=[local]) if (RESIDENT_WORKERS == 0 or profile) else \
            tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS, numWorkers=RESIDENT_WORKERS,
                devices=[local], tileCompletionTimeout=600)
        seg = tiling.TiledSegmenter(tiling.DeviceRaster(devImg, nB, bandRows, bandCols, numpy.uint16, yoff=y0, xoff=x0),
            range(1, nB + 1), tileInfo, wl['overlapSize'], centres, None, wl['fourConnected'],
            wl['minSegmentSize'], thr, False, cfg, timinghooks.Timers(), profile=profile)
        (maxSegId, hist) = seg.run(tiling.DeviceMosaicSink(devMosaic, bandCols, bandRows, yoff=y0, xoff=x0), comm)
        return (seg, maxSegId)

    def step_e2e():
        # the call a user makes: doTiledShepherdSegmentation, raster in (pinned) host memory, the
        # mosaic delivered into host memory, copies inside
        cfg = tiling.SegmentationConcurrencyConfig(concurrencyType=tiling.CONC_THREADS, numWorkers=E2E_WORKERS,
            devices=[local], tileCompletionTimeout=600)
        cfg.comm = comm
        sink = rasterfile.MemorySink(bandCols, bandRows, array=pinnedOut.array, yoff=y0, xoff=x0)
        src = rasterfile.MemoryRaster(img, yoff=y0, fullYsize=nR, xoff=x0, fullXsize=nC)
        res = tiling.doTiledShepherdSegmentation(src, sink, tileSize=wl['tileSize'], overlapSize=wl['overlapSize'],
            minSegmentSize=wl['minSegmentSize'], numClusters=wl['numClusters'], maxSpectralDiff=wl['maxSpectralDiff'],
            fourConnected=wl['fourConnected'], kmeansObj=km, concurrencyCfg=cfg, returnGDALDS=True)
        return (res, int(res.maxSegId))

    def exchange_ids(maxSegId):
        return 0      # (ids are global already: the sharded stitch numbers the mosaic as a whole)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    stepWalls = {}      # host wall time of every timed step of this rank

    def timed(stepFn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(steps):
            t0 = time.time()
            try:
                last = stepFn()
            except BaseException:
                # a rank that fails must not leave the others waiting in a collective
                import traceback
                traceback.print_exc()
                sys.stderr.flush()
                os._exit(1)
            exchange_ids(last[1])
            stepWalls.setdefault(getattr(stepFn, '__name__', '?'), []).append(round((time.time() - t0) * 1e3, 1))
            if args.verbose_steps and rank == 0:
                print('  step %s: %.1f ms' % (getattr(stepFn, '__name__', '?'), (time.time() - t0) * 1e3),
                    file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return (ms, last)

    # ---- warm-up -----------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
        step_e2e()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed: resident (value) with per-kernel events, then host-to-host (e2e) ------------------
    launches0 = sum(s.ctx.launch_count() for s in state.slots)
    kernelAgg = {}

    def resident_profiled():
        r = step_resident(profile=True)
        for (k, v) in r[0].kernelMs.items():
            ent = kernelAgg.setdefault(k, [0, 0.0])
            ent[0] += v[0]
            ent[1] += v[1]
        return r
    (msResident, lastRes) = timed(step_resident, args.steps)
    launchesResident = sum(s.ctx.launch_count() for s in state.slots) - launches0
    (msE2E, lastE2E) = timed(step_e2e, args.steps)
    # per-kernel durations for the roofline: the same steps once more on ONE stream with an event
    # pair around every launch (with several worker streams an event pair would also span the
    # other streams' kernels)
    step_resident(profile=True)       # (the first single-stream pass sizes its own scratch: not timed)
    (msProfiled, lastProf) = timed(resident_profiled, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if os.environ.get('SSG_TIMELINE') and rank == 0 and lastE2E[0].timeline:
        for (ms, what) in sorted(lastE2E[0].timeline):
            print('  %8.2f  %s' % (ms, what), file=sys.stderr)

    pixelsPerStep = nR * nC      # unique pixels of the mosaic (all ranks together)
    value = pixelsPerStep * args.steps / (msResident / 1e3) / 1e6
    e2eValue = pixelsPerStep * args.steps / (msE2E / 1e3) / 1e6
    segE2E = lastE2E[0]
    (h2dBytes, d2hBytes) = (int(segE2E.h2dBytes), int(segE2E.d2hBytes))
    if dist is not None:       # bytes of all ranks
        t = torch.tensor([h2dBytes, d2hBytes], dtype=torch.int64, device='cuda')
        dist.all_reduce(t)
        (h2dBytes, d2hBytes) = (int(t[0].item()), int(t[1].item()))

    # the e2e mosaic must be the resident mosaic (same labels, one went over PCIe)
    midRow = bandRows // 2
    check = numpy.empty((64, bandCols), dtype=numpy.uint32)
    ctx0.call('ssg_memcpy_d2h', _lib.ptr(check), devMosaic + midRow * bandCols * 4, check.nbytes)
    edge = wl['overlapSize'] if dist is not None else 0     # (the rim of a rank's window belongs to its neighbours)
    sameMosaic = bool(numpy.array_equal(check[:, edge:bandCols - edge], pinnedOut.array[midRow:midRow + 64, edge:bandCols - edge]))
    if dist is not None:
        t = torch.tensor([int(sameMosaic)], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        sameMosaic = bool(t.item())

    # ---- N ranks against one: the host-to-host mosaic of the last timed step, its windows put
    # together from all ranks, against rank 0 segmenting the same raster alone (untimed)
    vsOneRank = None
    if dist is not None and not args.no_parity:
        import zlib
        bounds = [None] * world
        dist.all_gather_object(bounds, (y0, y1, x0, x1))
        bandT = torch.from_numpy(img.view(numpy.int16)).cuda()
        full = None
        if rank == 0:
            full = torch.empty((nB, nR, nC), dtype=torch.int16, device='cuda')
            full[:, y0:y1, x0:x1] = bandT
            for r in range(1, world):
                (ry0, ry1, rx0, rx1) = bounds[r]
                tmp = torch.empty((nB, ry1 - ry0, rx1 - rx0), dtype=torch.int16, device='cuda')
                dist.recv(tmp.view(torch.uint8), r)      # (NCCL has no 16-bit integer type)
                full[:, ry0:ry1, rx0:rx1] = tmp
                del tmp
        else:
            dist.send(bandT.view(torch.uint8), 0)
        del bandT
        part = torch.zeros((nR, nC), dtype=torch.int32, device='cuda')
        for (cr, (x, y, xs, ys)) in tileInfo.tiles.items():
            if owner[cr] != rank:
                continue
            (top, bottom, left, right) = tiling.tileMargins(tileInfo, cr[0], cr[1], xs, ys, wl['overlapSize'])
            win = pinnedOut.array[y + top - y0:y + bottom - y0, x + left - x0:x + right - x0]
            part[y + top:y + bottom, x + left:x + right] = torch.from_numpy(
                numpy.ascontiguousarray(win).view(numpy.int32)).cuda()
        dist.reduce(part, 0)
        if rank == 0:
            torch.cuda.synchronize()
            one = torch.empty((nR, nC), dtype=torch.int32, device='cuda')
            seg1 = tiling.TiledSegmenter(tiling.DeviceRaster(full.data_ptr(), nB, nR, nC, numpy.uint16),
                range(1, nB + 1), tileInfo, wl['overlapSize'], centres, None, wl['fourConnected'],
                wl['minSegmentSize'], thr, False, tiling.SegmentationConcurrencyConfig(devices=[local]),
                timinghooks.Timers())
            (max1, hist1) = seg1.run(tiling.DeviceMosaicSink(one.data_ptr(), nC, nR))
            torch.cuda.synchronize()
            histN = getattr(getattr(lastE2E[0], 'outDs', None), 'hist', None)
            vsOneRank = {'equal': bool(torch.equal(part, one)) and int(max1) == int(lastE2E[1]),
                'segments': int(max1), 'segments_n_ranks': int(lastE2E[1]),
                'histogram_equal': None if histN is None else bool(numpy.array_equal(histN, numpy.asarray(hist1))),
                'crc32': '%08x' % (zlib.crc32(one.cpu().numpy()) & 0xffffffff),
                'what': 'e2e mosaic of the last timed step on %d ranks vs the same raster segmented by rank 0 alone' % world}
            del one, full
        del part
        torch.cuda.empty_cache()

    perRank = None
    if dist is not None:      # every rank's host timers of its last steps, for the scaling analysis
        mineT = {'resident': dict((k, round(v['total'] * 1e3, 1)) for (k, v) in
                    lastRes[0].timings.makeSummaryDict().items()),
                 'e2e': dict((k, round(v['total'] * 1e3, 1)) for (k, v) in
                    lastE2E[0].timings.makeSummaryDict().items()),
                 'step_wall_ms': stepWalls,
                 'tiles': len([cr for cr in tileInfo.tiles if owner[cr] == rank]),
                 'tile_mpix': round(sum(t[2] * t[3] for (cr, t) in tileInfo.tiles.items() if owner[cr] == rank) / 1e6, 1)}
        perRank = [None] * world
        dist.all_gather_object(perRank, mineT)
    if rank == 0:
        (peak, peakKind) = measured_peaks()
        # roofline of the dominant kernel, from the events recorded around every launch
        dom = max(kernelAgg.items(), key=lambda kv: kv[1][1]) if kernelAgg else (None, [0, 0.0])
        (domName, (domCount, domMs)) = dom
        tilePixels = sum(t[2] * t[3] for (cr, t) in tileInfo.tiles.items()
            if world == 1 or owner[cr] == rank)     # the kernel times below are rank 0's
        bpp = algorithmic_bytes_per_pixel(domName, nB)
        roof = {'bound': 'hbm', 'kernel': domName, 'achieved': None, 'peak': peak, 'unit': 'GB/s',
            'frac': None, 'traffic': None, 'peak_source': peakKind}
        if bpp is not None and domCount > 0 and domMs > 0:
            # one launch per tile: algorithmic bytes per launch = bytes/pixel x mean tile pixels
            launchesPerStep = domCount / args.steps
            bytesPerLaunch = bpp * tilePixels / max(1.0, launchesPerStep)
            avgMs = domMs / domCount
            roof['achieved'] = bytesPerLaunch / (avgMs / 1e3) / 1e9
            roof['frac'] = roof['achieved'] / peak
            roof['bytes_per_pixel'] = bpp
            if domName in NCU_DRAM_BYTES_PER_PIXEL:
                roof['traffic'] = NCU_DRAM_BYTES_PER_PIXEL[domName] * tilePixels / max(1.0, launchesPerStep)
                roof['traffic_source'] = ('ncu --set full on a 4096x4096x4 tile (profiles/r2_tile4096_ncu_full.md): '
                    '%.1f DRAM bytes per pixel, scaled to the mean tile of a launch' % NCU_DRAM_BYTES_PER_PIXEL[domName])
            roof['avg_launch_ms'] = avgMs
            roof['launches_per_step'] = launchesPerStep
            roof['share_of_step'] = domMs / msProfiled
            roof['timed_in'] = ('%d single-stream steps after the timed region, event pair per launch '
                '(%.1f ms per step)' % (args.steps, msProfiled / args.steps))
        kernels = dict((k, {'launches': v[0], 'ms': round(v[1], 3)}) for (k, v) in sorted(kernelAgg.items(),
            key=lambda kv: -kv[1][1])[:24])
        # secondary rooflines of the two bandwidth kernels the north star names
        extra = {}
        for name in ('k_assign', 'k_assign_grid', 'k_ccl_local', 'k_pixel_pass', 'k_gather_ids', 'k_apply_lut_extents'):
            if name in kernelAgg and kernelAgg[name][1] > 0:
                b = algorithmic_bytes_per_pixel(name, nB)
                gbs = b * tilePixels * args.steps / (kernelAgg[name][1] / 1e3) / 1e9
                extra[name] = {'achieved': gbs, 'frac': gbs / peak, 'bytes_per_pixel': b}

        # ---- the CPU path beside it (rank 0, N = 1 only): the unmodified reference (numba) in a
        # child process of this script, and the C port in threads here
        cpu = None
        cpuPort = None
        ratio = tile_pixel_ratio(tileInfo, nR, nC)
        if args.gpus == 1 and not args.no_cpu_baseline:
            threads = min(os.cpu_count() or 1, 64)
            tile = 2048 if args.quick else 4096
            (p, s) = cpu_port_sample(wl, img, centres, threads, tile=tile)
            cpuPort = {'value': p / ratio / s / 1e6, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                'sample': '%d tiles of %dx%dx%d cut from the workload raster, one per thread (%.1f s); unique-pixel '
                    'credit = tile pixels / %.3f' % (threads, tile, tile, nB, s, ratio)}
            cpu = cpuPort
            try:
                cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '2', '--warmup', '1']
                if args.quick:
                    cmd.append('--quick')
                out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, check=True)
                ref = json.loads(out.stdout.decode().strip().splitlines()[-1])['cpu_baseline']
                if ref.get('kind') == 'reference':
                    cpu = ref
            except Exception as e:
                print('numba reference sample failed: %r' % (e,), file=sys.stderr)

        # ---- parity of the BENCHED mosaic: the host-to-host result of the last timed step against
        # the CPU oracle run on the same raster (every tile through oracle.doShepherdSegmentation on
        # the host threads, then oracle.stitchTiles), outside the timed regions
        parity = None
        mosaicCrc = None
        if args.gpus == 1 and not args.no_parity:
            import zlib
            from oracle import oracle
            oracle.lib()
            t0 = time.time()
            kmO = KM(centres)
            keys = sorted(tileInfo.tiles.keys(), key=lambda cr: -tileInfo.tiles[cr][2] * tileInfo.tiles[cr][3])
            segs = {}

            def oracleTile(cr):
                (x, y, xs, ys) = tileInfo.tiles[cr]
                sub = numpy.ascontiguousarray(img[:, y:y + ys, x:x + xs])
                segs[cr] = oracle.doShepherdSegmentation(sub, minSegmentSize=wl['minSegmentSize'],
                    maxSpectralDiff=wl['maxSpectralDiff'], fourConnected=wl['fourConnected'], kmeansObj=kmO).segimg
            ths = [threading.Thread(target=oracleTile, args=(cr,)) for cr in keys]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            oti = oracle.getTilesForFile(nC, nR, wl['tileSize'], wl['overlapSize'])
            (mosaicO, maxO, histO) = oracle.stitchTiles(segs, oti, nC, nR, wl['overlapSize'])
            same = bool(numpy.array_equal(mosaicO, pinnedOut.array)) and int(maxO) == int(lastE2E[1])
            mosaicCrc = '%08x' % (zlib.crc32(pinnedOut.array) & 0xffffffff)
            parity = {'equal': same, 'segments': int(maxO), 'oracle_crc32': '%08x' % (zlib.crc32(mosaicO) & 0xffffffff),
                'seconds': round(time.time() - t0, 1),
                'what': 'e2e mosaic of the last timed step vs oracle (per-tile C port + stitch) on the same raster'}
            del mosaicO, segs
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': msResident / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if wl.get('strong') else 'weak', 'vs_baseline': None, 'dtype': 'u16', 'data': 'synthetic',
            'config': workload_config(wl, args.gpus, (nR, nC), tileInfo),
            'e2e': {'value': e2eValue, 'unit': UNIT, 'ms_per_step': msE2E / args.steps,
                'h2d_bytes_per_step': h2dBytes, 'd2h_bytes_per_step': d2hBytes,
                'workers': E2E_WORKERS, 'same_labels_as_resident': sameMosaic},
            'gpu_launches': int(launchesResident),
            'roofline': roof, 'roofline_other': extra, 'kernels': kernels,
            'cpu_baseline': cpu, 'cpu_baseline_port': cpuPort, 'clocks': clocks,
            'parity_vs_oracle': None if parity is None else parity['equal'], 'parity': parity,
            'mosaic_crc32': mosaicCrc, 'parity_vs_one_rank': None if vsOneRank is None else vsOneRank['equal'],
            'one_rank': vsOneRank,
            'segments_per_scene': int(lastRes[1]),
            'stage_ms_per_step': dict((k, round(v, 3)) for (k, v) in lastProf[0].stageMs.items()),
            # the stages against the HBM roofline with SURVEY 8(d)'s per-pixel figures: assign reads the
            # bands and writes a label, clump reads and writes a label, the two eliminate stages read
            # the bands once and read + write the labels
            'roofline_stages': dict((k, {'bytes_per_pixel': b, 'achieved': b * tilePixels / (lastProf[0].stageMs[k] / 1e3) / 1e9,
                    'frac': b * tilePixels / (lastProf[0].stageMs[k] / 1e3) / 1e9 / peak})
                for (k, b) in (('assign', 2 * nB + 4), ('clump', 8), ('single', 2 * nB + 8), ('small', 2 * nB + 8))
                if lastProf[0].stageMs.get(k)),
            'per_rank': perRank,
            'host_ms_last_step': {
                'resident': dict((k, round(v['total'] * 1e3, 2)) for (k, v) in
                    lastRes[0].timings.makeSummaryDict().items()),
                'e2e': dict((k, round(v['total'] * 1e3, 2)) for (k, v) in
                    lastE2E[0].timings.makeSummaryDict().items())},
        }
        print(json.dumps(line), flush=True)

    ctx0.dev_free(devImg)
    ctx0.dev_free(devMosaic)
    tiling.releaseGpuState()
    pinnedImg.free()
    pinnedOut.free()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--quick', action='store_true', help='2048-pixel raster and tiles (smoke runs only)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true', help='skip the oracle check of the benched mosaic')
    ap.add_argument('--port-only', action='store_true', help='reference arm: time the C port even if baseline/_ref exists')
    ap.add_argument('--verbose-steps', action='store_true', help='print every timed step (stderr)')
    ap.add_argument('--workload', default='c2', choices=['c2', 'c4_40000'],
        help='c2: the 10980x10980x4 scene (N > 1: a mosaic of N scenes, weak scaling); c4_40000: BASELINE '
             'config 4, one 40000x40000x4 mosaic for every N (strong scaling)')
    args = ap.parse_args()
    wl = dict(WORKLOAD)
    if args.quick:
        wl.update({'name': 'quick_2700x2700x4', 'rows': 2700, 'cols': 2700, 'tileSize': 1024, 'overlapSize': 256})
    if args.workload == 'c4_40000':
        wl.update({'name': 'mosaic_40000x40000x4_u16_tiled', 'rows': 40000, 'cols': 40000, 'strong': True})
        if args.quick:
            wl.update({'name': 'quick_mosaic_6000x6000x4', 'rows': 6000, 'cols': 6000})
    if args.impl == 'reference':
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == '__main__':
    main()
