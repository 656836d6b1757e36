"""
One mosaic, its tiles sharded over the GPUs of one box (one process per GPU, torch.distributed).

The path shards by tile: every tile is segmented independently given the shared cluster
centres (shepseg.doShepherdSegmentation per tile, tiling.py:1446/1586).  What the reference does
sequentially afterwards (stitchTiles, tiling.py:950-1064: tile after tile in row-major order,
carrying the running maxSegId and the recoded overlap strips of the finished neighbours) needs
three small exchanges when the tiles live on different ranks:

  strips   the LOCAL labels of an upper / left neighbour under a tile's overlap, when that
           neighbour lives on another rank (device-to-device send/recv, NCCL over NVLink);
  counts   one integer per tile, the highest rank among its self-numbered segments inside its
           trimmed window: the id offset of tile t is the sum over the tiles before it (one
           all-gather).  That is what the reference's "maxSegId = max(maxSegId, trimmed.max())"
           (tiling.py:1042-1043) amounts to unless a window holds an inherited id above the
           running maximum; the recurrence is checked on every tile after the resolve and the
           ranks fall back to the sequential order over all tables if it fails anywhere, so the
           result is the reference's in every case;
  tables   the per-segment tables (rank, flags, votes) of the tiles whose final ids a tile on
           another rank has to look up: a crossing segment takes the final id of the neighbour
           segment it overlaps most (recodeSharedSegments, tiling.py:1128-1203), and that id may
           in turn be inherited from the neighbour's neighbour.  Look-ups are lazy and per
           entry, so no rank replays another rank's tiles.

Everything here is host logic on numpy arrays; the device work is behind the `ops` object the
caller hands in (tiling.TiledSegmenter for the GPU, plain numpy in the CPU tests).
"""
import io

import numpy

from . import _lib

KEY_FLAGS = _lib.SEG_KEYTOP | _lib.SEG_KEYLEFT


def rowMajor(tileInfo):
    return sorted(tileInfo.tiles.keys(), key=lambda cr: (cr[1], cr[0]))


# cost of segmenting a tile, in arbitrary units: a fixed part (about a hundred dependent phases of
# the merge kernel, whatever the tile's size) plus a part per pixel; measured on B200 as
# 1.2 ms + 0.13 ms per megapixel
TILE_COST_FIXED = 1.2
TILE_COST_PER_MPIX = 0.13


def partitionTiles(tileInfo, world):
    """
    owner rank of every tile: contiguous chunks of the row-major tile list (tiling.py:892-893)
    with as equal a share of the segmentation cost as the tile boundaries allow.  Contiguous
    chunks keep most neighbours on the same GPU.
    """
    order = rowMajor(tileInfo)
    cost = numpy.array([TILE_COST_FIXED + TILE_COST_PER_MPIX * tileInfo.tiles[cr][2] * tileInfo.tiles[cr][3] / 1e6
        for cr in order], dtype=numpy.float64)
    cum = numpy.cumsum(cost)
    total = cum[-1]
    owner = {}
    for (i, cr) in enumerate(order):
        mid = cum[i] - cost[i] / 2.0          # the rank whose share holds the tile's midpoint
        owner[cr] = min(world - 1, int(mid * world / total))
    # a rank must not be skipped: make the assignment monotone and gap-free
    last = 0
    for cr in order:
        r = owner[cr]
        if r > last + 1:
            r = last + 1
        owner[cr] = r
        last = r
    return owner


class TileTable(object):
    """The host copy of what ssg_tile_tables_device computed for one tile."""
    def __init__(self, maxId, countNew, rank, flags, pairKeys, pairCounts):
        self.maxId = int(maxId)
        self.countNew = int(countNew)
        self.rank = numpy.ascontiguousarray(rank, dtype=numpy.uint32)
        self.flags = numpy.ascontiguousarray(flags, dtype=numpy.uint8)
        self.pairKeys = numpy.ascontiguousarray(pairKeys, dtype=numpy.uint64)
        self.pairCounts = numpy.ascontiguousarray(pairCounts, dtype=numpy.uint32)
        self._split = None

    @property
    def maxRankInTrim(self):
        sel = (self.flags & (_lib.SEG_NUMBERED | _lib.SEG_INTRIM)) == (_lib.SEG_NUMBERED | _lib.SEG_INTRIM)
        return int(self.rank[sel].max()) if sel.any() else 0

    @property
    def maxLabelInTrim(self):
        sel = numpy.flatnonzero(self.flags & _lib.SEG_INTRIM)
        return int(sel[-1]) if len(sel) else 0

    def prepare(self):
        """Everything of the tile's lut that does not depend on the id offset, so that it can be
        worked out as soon as the tables exist (while other tiles are still being segmented):
        (rank where numbered else 0, 1 where numbered else 0, positions of the crossing segments)"""
        if getattr(self, '_prepared', None) is None:
            numbered = ((self.flags & _lib.SEG_NUMBERED) != 0)
            rel = numpy.where(numbered, self.rank, numpy.uint32(0)).astype(numpy.uint32)
            crossing = numpy.flatnonzero((self.flags & KEY_FLAGS) != 0)
            self._prepared = (rel, numbered.astype(numpy.uint32), crossing)
            self.crossingInTrim = crossing[(self.flags[crossing] & _lib.SEG_INTRIM) != 0]
            self.cachedMaxRankInTrim = self.maxRankInTrim
            self.pairs()
        return self._prepared

    def pairs(self):
        """(isLeft, segment, neighbour label, count) of the votes, decoded once."""
        if self._split is None:
            k = self.pairKeys
            self._split = ((k >> numpy.uint64(63)) != 0,
                ((k >> numpy.uint64(32)) & numpy.uint64(0x7FFFFFFF)).astype(numpy.int64),
                (k & numpy.uint64(0xFFFFFFFF)).astype(numpy.int64),
                self.pairCounts.astype(numpy.int64))
        return self._split


def packTables(tables):
    """{tile: TileTable} -> one uint8 array (for an all-gather of byte buffers)."""
    items = sorted(tables.items())
    total = 8
    for (cr, tb) in items:
        total += 24 + 32 + 8 * len(tb.pairKeys) + 4 * len(tb.rank) + 4 * len(tb.pairCounts) + len(tb.flags)
        total += (-total) % 8
    out = numpy.zeros(total, dtype=numpy.uint8)
    out[:8].view(numpy.int64)[0] = len(items)
    o = 8
    for (cr, tb) in items:
        body = 32 + 8 * len(tb.pairKeys) + 4 * len(tb.rank) + 4 * len(tb.pairCounts) + len(tb.flags)
        size = body + ((-(o + 24 + body)) % 8)
        out[o:o + 24].view(numpy.int64)[:] = (cr[0], cr[1], size)
        o += 24
        out[o:o + 32].view(numpy.int64)[:] = (tb.maxId, tb.countNew, len(tb.rank), len(tb.pairKeys))
        q = o + 32
        for a in (tb.pairKeys, tb.rank, tb.pairCounts, tb.flags):      # widest first: aligned
            nb = a.nbytes
            out[q:q + nb] = a.view(numpy.uint8)
            q += nb
        o += size
    return out


def unpackTables(b):
    b = numpy.frombuffer(b, dtype=numpy.uint8) if not isinstance(b, numpy.ndarray) else b
    n = int(b[:8].view(numpy.int64)[0])
    o = 8
    out = {}
    for _ in range(n):
        (c, r, size) = (int(v) for v in b[o:o + 24].view(numpy.int64))
        o += 24
        (maxId, countNew, nSeg, nPairs) = (int(v) for v in b[o:o + 32].view(numpy.int64))
        q = o + 32
        pairKeys = b[q:q + 8 * nPairs].view(numpy.uint64)
        q += 8 * nPairs
        rank = b[q:q + 4 * nSeg].view(numpy.uint32)
        q += 4 * nSeg
        pairCounts = b[q:q + 4 * nPairs].view(numpy.uint32)
        q += 4 * nPairs
        flags = b[q:q + nSeg]
        out[(c, r)] = TileTable(maxId, countNew, rank, flags, pairKeys, pairCounts)
        o += size
    return out


class MissingTable(Exception):
    """a look-up reached a tile whose table this rank does not hold"""


class LazyResolver(object):
    """
    Final ids from per-tile tables when the id offsets are known up front (offset of a tile = sum
    of countNew over the tiles before it).  lut entries are computed on demand:
      numbered segment           offset + rank                       (tiling.py:1264-1267)
      crossing segment           the mode of the final ids of the neighbour's labels under it,
                                 smallest id on ties, top overlap first, left overriding
                                 (tiling.py:1107-1121, 1194)
      anything else, and 0       0 (the segment belongs to a neighbour, tiling.py:1241)
    """
    UNKNOWN = numpy.uint32(0xFFFFFFFF)

    def __init__(self, tables, offsets, simple=False):
        self.tables = tables
        self.offsets = offsets
        self.simple = simple
        self.luts = {}
        self.misses = 0         # look-ups that could not be answered yet
        self.remote = {}        # tile of another rank -> (labels ascending, their final ids)
        self.requests = {}      # tile of another rank -> [label arrays] still to be asked for

    def _lutOf(self, cr):
        """the tile's lut with everything that needs no neighbour filled in (one vectorised pass);
        crossing segments start out unknown"""
        if cr not in self.luts:
            if cr not in self.tables:
                raise MissingTable(cr)
            tb = self.tables[cr]
            off = numpy.uint32(self.offsets[cr])
            if self.simple:
                lut = numpy.arange(tb.maxId + 1, dtype=numpy.uint32) + off
                lut[0] = 0
            else:
                (rel, isNumbered, crossing) = tb.prepare()
                lut = rel + isNumbered * off
                lut[crossing] = self.UNKNOWN
            self.luts[cr] = lut
        return self.luts[cr]

    def remoteIds(self, cr, labels):
        """final ids of labels of a tile owned by another rank, from the owner's answers so far;
        what is not known yet is queued in self.requests.  Returns (ids, complete)."""
        labels = numpy.asarray(labels, dtype=numpy.int64)
        out = numpy.zeros(len(labels), dtype=numpy.uint32)
        found = numpy.zeros(len(labels), dtype=bool)
        if cr in self.remote:
            (known, vals) = self.remote[cr]
            pos = numpy.searchsorted(known, labels)
            pos[pos >= len(known)] = 0
            found = known[pos] == labels
            out[found] = vals[pos[found]]
        if not found.all():
            self.requests.setdefault(cr, []).append(labels[~found])
            return (out, False)
        return (out, True)

    def addRemote(self, cr, labels, vals):
        if cr in self.remote:
            labels = numpy.concatenate([self.remote[cr][0], labels])
            vals = numpy.concatenate([self.remote[cr][1], vals])
        order = numpy.argsort(labels, kind='stable')
        self.remote[cr] = (labels[order].astype(numpy.int64), vals[order].astype(numpy.uint32))

    def answer(self, cr, labels):
        """what an owner can say about labels of its tile cr right now: (labels, final ids) of
        those that are settled"""
        labels = numpy.asarray(labels, dtype=numpy.int64)
        self.finalIds(cr, labels)
        vals = self.luts[cr][labels]
        settled = vals != self.UNKNOWN
        return (labels[settled], vals[settled])

    def finalIds(self, cr, labels):
        """final ids of the given local labels of tile cr (int64 array).  Entries whose value
        hangs on a table that is not here stay unknown (and come back as 0): the tile is noted in
        self.missing, to be fetched before the next call."""
        lut = self._lutOf(cr)
        labels = numpy.asarray(labels, dtype=numpy.int64)
        out = lut[labels]
        unknown = out == self.UNKNOWN
        if unknown.any():
            self._resolveCrossing(cr, lut, labels[unknown])
            out = lut[labels]
        return numpy.where(out == self.UNKNOWN, 0, out).astype(numpy.uint32)

    def _resolveCrossing(self, cr, lut, labels):
        """A crossing segment takes the vote of the left overlap if it crosses that one (the left
        recode is applied after the top one and overrides it, tiling.py:1107-1121), otherwise the
        vote of the top overlap."""
        from .tiling import _modeByKey
        tb = self.tables[cr]
        (isLeft, segs, nbr, counts) = tb.pairs()
        wanted = numpy.zeros(tb.maxId + 1, dtype=bool)
        wanted[labels] = True
        keyLeft = (tb.flags & _lib.SEG_KEYLEFT) != 0
        for (which, pairSide, nb) in ((wanted & ~keyLeft, ~isLeft, (cr[0], cr[1] - 1)),
                (wanted & keyLeft, isLeft, (cr[0] - 1, cr[1]))):
            sel = pairSide & which[segs]
            if not sel.any():
                lut[which] = 0
                continue
            if nb not in self.tables:
                # a tile of another rank: its owner answers (entries stay open until then)
                (mapped, complete) = self.remoteIds(nb, nbr[sel])
                if not complete:
                    self.misses += 1
                    continue
                mapped = mapped.astype(numpy.int64)
            else:
                before = self.misses
                mapped = self.finalIds(nb, nbr[sel]).astype(numpy.int64)
                if self.misses > before:
                    continue          # the neighbour's own answer is still open somewhere below
            lut[which] = 0
            (k, mode) = _modeByKey(segs[sel], mapped, counts[sel])
            lut[k] = mode.astype(numpy.uint32)

    def settle(self, cr):
        """work on every open entry of own tile cr; True when none is left open"""
        lut = self._lutOf(cr)
        if self.simple:
            return True
        crossing = self.tables[cr].prepare()[2]
        unknown = crossing[lut[crossing] == self.UNKNOWN]
        if len(unknown) > 0:
            self._resolveCrossing(cr, lut, unknown)
            return not (lut[unknown] == self.UNKNOWN).any()
        return True

    def fullLut(self, cr):
        self.settle(cr)
        lut = self._lutOf(cr)
        return numpy.where(lut == self.UNKNOWN, 0, lut).astype(numpy.uint32)


def sequentialResolve(order, tables, simple=False):
    """The reference's order, tile after tile (tiling.resolveTile); returns (luts, offsets, maxSegId)."""
    from .tiling import resolveTile
    luts = {}
    offsets = {}
    offset = 0
    for cr in order:
        tb = tables[cr]
        offsets[cr] = offset
        t = _lib.TileTables()
        t.maxId = tb.maxId
        t.countNew = tb.countNew
        t.numPairs = len(tb.pairKeys)
        (lut, trimmedMax) = resolveTile(t, tb.rank, tb.flags, tb.pairKeys, tb.pairCounts, offset,
            luts.get((cr[0], cr[1] - 1)), luts.get((cr[0] - 1, cr[1])), simple)
        luts[cr] = lut
        offset = max(offset, trimmedMax)
    return (luts, offsets, offset)


class LocalComm(object):
    """world of one (and the interface the real communicators implement)"""
    rank = 0
    world = 1

    def allgatherInts(self, values):
        """list of ints from every rank -> list (per rank) of lists"""
        return [list(values)]

    def allgatherBytes(self, b):
        """uint8 array (or bytes) from every rank -> list of uint8 arrays"""
        return [numpy.frombuffer(b, dtype=numpy.uint8) if not isinstance(b, numpy.ndarray) else b]

    def allgatherArray(self, a):
        """int64 array from every rank -> list of int64 arrays"""
        return [numpy.asarray(a, dtype=numpy.int64)]

    def exchange(self, sends, recvs):
        """sends: [(dstRank, tensor)], recvs: [(srcRank, tensor)] in matching order per rank pair"""
        assert not sends and not recvs

    def allreduceSum(self, array):
        return array


class TorchComm(object):
    """torch.distributed process group (nccl on the GPUs, gloo in the CPU tests)"""
    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch = torch
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.device = device if device is not None else torch.device('cpu')

    def _t(self, a):
        return self.torch.from_numpy(numpy.ascontiguousarray(a)).to(self.device)

    def allgatherInts(self, values):
        n = self._t(numpy.array([len(values)], dtype=numpy.int64))
        sizes = [self.torch.zeros_like(n) for _ in range(self.world)]
        self.dist.all_gather(sizes, n)
        sizes = [int(s.item()) for s in sizes]
        m = max(sizes) if sizes else 0
        mine = numpy.zeros(max(m, 1), dtype=numpy.int64)
        mine[:len(values)] = values
        mine = self._t(mine)
        out = [self.torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(out, mine)
        return [o.cpu().numpy()[:sizes[i]].tolist() for (i, o) in enumerate(out)]

    def allgatherBytes(self, b):
        if not isinstance(b, numpy.ndarray):
            b = numpy.frombuffer(b, dtype=numpy.uint8)
        sizes = [s[0] for s in self.allgatherInts([len(b)])]
        m = max(max(sizes), 1)
        mine = numpy.zeros(m, dtype=numpy.uint8)
        mine[:len(b)] = b
        mine = self._t(mine)
        out = [self.torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(out, mine)
        return [o.cpu().numpy()[:sizes[i]] for (i, o) in enumerate(out)]

    def allgatherArray(self, a, quick=1 << 15):
        """int64 array from every rank -> list of int64 arrays.  One collective when every rank's
        array fits `quick` entries (slot 0 carries the length), otherwise the sizes first."""
        a = numpy.ascontiguousarray(a, dtype=numpy.int64)
        torch = self.torch
        if not hasattr(self, '_quickBufs') or self._quickBufs[0].numel() != quick + 1:
            self._quickBufs = (torch.zeros(quick + 1, dtype=torch.int64, device=self.device),
                torch.zeros(self.world * (quick + 1), dtype=torch.int64, device=self.device))
        (mine, everyone) = self._quickBufs
        host = numpy.zeros(quick + 1, dtype=numpy.int64)
        host[0] = len(a)
        if len(a) <= quick:
            host[1:1 + len(a)] = a
        mine.copy_(torch.from_numpy(host))
        self.dist.all_gather_into_tensor(everyone, mine)
        got = everyone.cpu().numpy().reshape(self.world, quick + 1)
        sizes = got[:, 0]
        if (sizes <= quick).all():
            return [got[r, 1:1 + int(sizes[r])].copy() for r in range(self.world)]
        return [b.view(numpy.int64) for b in self.allgatherBytes(a.view(numpy.uint8))]

    def exchange(self, sends, recvs):
        ops = []
        for (dst, t) in sends:
            ops.append(self.dist.P2POp(self.dist.isend, t, dst))
        for (src, t) in recvs:
            ops.append(self.dist.P2POp(self.dist.irecv, t, src))
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        if self.device.type == 'cuda':
            self.torch.cuda.synchronize(self.device)

    def allreduceSum(self, array):
        t = self._t(array)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()


class ShardedStitch(object):
    """
    The stitch of a mosaic whose tiles are spread over the ranks of `comm`.  The caller has
    segmented its own tiles and provides `ops`:

      ops.sendStrip(cr, which)            the LOCAL labels of tile cr's bottom ('bottom': last
                                          overlap rows) or right ('right': last overlap columns)
                                          strip as a contiguous tensor to send
      ops.recvStrip(cr, which, shape)     a tensor to receive such a strip of a remote tile into
      ops.tables(cr, top, left)           TileTable of own tile cr; top / left are None (no such
                                          neighbour), 'local' (the neighbour is an own tile) or
                                          the tensor received for it
      ops.apply(cr, lut, table)           write lut[tile] over the trimmed window of own tile cr
    """
    def __init__(self, tileInfo, overlapSize, simple, comm, timings=None):
        self.timings = timings
        self.tileInfo = tileInfo
        self.overlap = int(overlapSize)
        self.simple = simple
        self.comm = comm
        self.order = rowMajor(tileInfo)
        self.owner = partitionTiles(tileInfo, comm.world)
        self.mine = [cr for cr in self.order if self.owner[cr] == comm.rank]
        self.usedFallback = False
        self.forceSequential = False    # (tests) take the fall-back even if the check passes

    def neighbours(self, cr):
        (c, r) = cr
        up = (c, r - 1) if r > 0 else None
        left = (c - 1, r) if c > 0 else None
        return (up, left)

    def stripPlan(self):
        """(sends, recvs) of this rank as lists of (peerRank, tile, which, shape), in the
        row-major order of the receiving tile (both sides enumerate the same order)."""
        sends = []
        recvs = []
        ov = self.overlap
        if self.simple:
            return (sends, recvs)
        for cr in self.order:
            (xpos, ypos, xsize, ysize) = self.tileInfo.tiles[cr]
            (up, left) = self.neighbours(cr)
            for (nb, which) in ((up, 'bottom'), (left, 'right')):
                if nb is None or self.owner[nb] == self.owner[cr]:
                    continue
                (nx, ny, nxs, nys) = self.tileInfo.tiles[nb]
                shape = (ov, nxs) if which == 'bottom' else (nys, ov)
                if self.owner[nb] == self.comm.rank:
                    sends.append((self.owner[cr], nb, which, shape))
                if self.owner[cr] == self.comm.rank:
                    recvs.append((self.owner[nb], nb, which, shape))
        return (sends, recvs)

    def _timed(self, name):
        import contextlib
        return self.timings.interval(name) if self.timings is not None else contextlib.nullcontext()

    def run(self, ops):
        """returns (maxSegId, offsets of all tiles, luts of own tiles)"""
        with self._timed('stitch_strips'):
            received = self._exchangeStrips(ops)
        with self._timed('stitch_owntables'):
            tables = self._ownTables(ops, received)
        with self._timed('stitch_offsets'):
            (steps, offsets, maxSegId) = self._offsets(tables)
        with self._timed('stitch_resolve'):
            (luts, offsets, maxSegId) = self._resolve(tables, steps, offsets, maxSegId)
        with self._timed('stitch_apply'):
            for cr in self.mine:
                ops.apply(cr, luts[cr], tables[cr])
        return (maxSegId, offsets, luts)

    def _exchangeStrips(self, ops):
        comm = self.comm
        # 1. strips of remote neighbours
        (sends, recvs) = self.stripPlan()
        received = {}
        sendList = [(peer, ops.sendStrip(cr, which)) for (peer, cr, which, shape) in sends]
        recvList = []
        for (peer, cr, which, shape) in recvs:
            t = ops.recvStrip(cr, which, shape)
            received[(cr, which)] = t
            recvList.append((peer, t))
        comm.exchange(sendList, recvList)
        return received

    def _ownTables(self, ops, received):
        comm = self.comm
        # 2. tables of own tiles
        tables = {}
        for cr in self.mine:
            (up, left) = self.neighbours(cr)
            top = lf = None
            if not self.simple:
                if up is not None:
                    top = 'local' if self.owner[up] == comm.rank else received[(up, 'bottom')]
                if left is not None:
                    lf = 'local' if self.owner[left] == comm.rank else received[(left, 'right')]
            tables[cr] = ops.tables(cr, top, lf)
        return tables

    def _offsets(self, tables):
        comm = self.comm
        # 3. offsets.  The reference moves the running maximum on tile after tile:
        # maxSegId = max(maxSegId, trimmed.max()) (tiling.py:1042-1043).  Inside the trimmed
        # window a tile has its own numbered segments (offset + rank) and segments recoded to ids
        # of EARLIER tiles; unless one of those carries an id above the running maximum (a
        # neighbour's segment that was numbered without having a pixel in its own trimmed
        # window), the maximum moves on by the highest rank present in the window.  Take that as
        # the hypothesis -- offsets = running sum of one integer per tile, known to every rank
        # after one all-gather -- resolve, and check the recurrence on every tile afterwards.
        mineInts = []
        for cr in self.mine:
            tb = tables[cr]
            step = tb.maxLabelInTrim if self.simple else (tb.prepare() and tb.cachedMaxRankInTrim)
            mineInts += [cr[0], cr[1], step]
        steps = {}
        for vals in comm.allgatherArray(numpy.array(mineInts, dtype=numpy.int64)):
            for i in range(0, len(vals), 3):
                steps[(int(vals[i]), int(vals[i + 1]))] = int(vals[i + 2])
        offsets = {}
        offset = 0
        for cr in self.order:
            offsets[cr] = offset
            offset += steps[cr]
        maxSegId = offset
        return (steps, offsets, maxSegId)

    def _resolve(self, tables, steps, offsets, maxSegId):
        comm = self.comm
        # 4. final ids of own tiles.  A crossing segment takes the final id of the neighbour
        # segment it overlaps most; when that neighbour tile lives on another rank its OWNER is
        # asked for the final ids of the labels in question (a few thousand per boundary tile),
        # and answers what it has settled.  An id inherited through several tiles around a
        # corner takes one more round per hop.
        luts = {}
        resolver = LazyResolver(dict(tables), offsets, self.simple)
        for _round in range(len(self.order) + 3):
            resolver.requests = {}
            for cr in self.mine:          # (entries settled in an earlier round are kept)
                resolver.settle(cr)
            req = []
            for (cr, parts) in sorted(resolver.requests.items()):
                labels = numpy.unique(numpy.concatenate(parts))
                req += [numpy.array([cr[0], cr[1], len(labels)], dtype=numpy.int64), labels]
            allReq = comm.allgatherArray(numpy.concatenate(req) if req else numpy.zeros(0, numpy.int64))
            if not any(len(a) for a in allReq):
                break
            ans = []
            for a in allReq:
                o = 0
                while o < len(a):
                    (c, r, n) = (int(v) for v in a[o:o + 3])
                    labels = a[o + 3:o + 3 + n]
                    o += 3 + n
                    if self.owner[(c, r)] == comm.rank:
                        (lab, vals) = resolver.answer((c, r), labels)
                        ans += [numpy.array([c, r, len(lab)], dtype=numpy.int64), lab, vals.astype(numpy.int64)]
            for a in comm.allgatherArray(numpy.concatenate(ans) if ans else numpy.zeros(0, numpy.int64)):
                o = 0
                while o < len(a):
                    (c, r, n) = (int(v) for v in a[o:o + 3])
                    if (c, r) in resolver.requests:
                        resolver.addRemote((c, r), a[o + 3:o + 3 + n], a[o + 3 + n:o + 3 + 2 * n])
                    o += 3 + 2 * n
        else:
            raise RuntimeError('sharded stitch: look-ups across ranks did not settle')
        for cr in self.mine:
            luts[cr] = resolver.luts[cr]          # nothing is open any more
        # the check of the hypothesis
        ok = 1
        for cr in self.mine:
            tb = tables[cr]
            if self.simple:
                trimmedMax = offsets[cr] + tb.maxLabelInTrim if tb.maxLabelInTrim else 0
            else:
                tb.prepare()
                trimmedMax = offsets[cr] + tb.cachedMaxRankInTrim if tb.cachedMaxRankInTrim else 0
                if len(tb.crossingInTrim):
                    trimmedMax = max(trimmedMax, int(luts[cr][tb.crossingInTrim].max()))
            if max(offsets[cr], trimmedMax) != offsets[cr] + steps[cr] or self.forceSequential:
                ok = 0
        if min(int(v[0]) for v in comm.allgatherArray(numpy.array([ok], dtype=numpy.int64))) == 0:
            # some window holds an inherited id above the running maximum: replay the
            # reference's sequential order over all tables, on every rank
            self.usedFallback = True
            known = {}
            for b in comm.allgatherBytes(packTables(tables)):
                known.update(unpackTables(b))
            (allLuts, offsets, maxSegId) = sequentialResolve(self.order, known, self.simple)
            luts = dict((cr, allLuts[cr]) for cr in self.mine)

        return (luts, offsets, maxSegId)
