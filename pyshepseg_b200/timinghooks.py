"""
Named wall-clock interval timers, as returned in TiledSegmentationResult.timings.

Same observable surface as pyshepseg.timinghooks.Timers (timinghooks.py:18-160): an
`interval(name)` context manager, thread safe, `getDurationsForName`, `merge`,
`makeSummaryDict`.  The interval names used by the tiled driver are the reference's
(walltime, spectralclusters, startworkers, reading, segmentation, stitchtiles).
"""
import contextlib
import threading
import time

import numpy


class Timers(object):
    def __init__(self, pairs=None, withLock=True):
        self.pairs = {} if pairs is None else pairs
        self.lock = threading.Lock() if withLock else None

    @contextlib.contextmanager
    def interval(self, intervalName):
        start = time.time()
        try:
            yield
        finally:
            end = time.time()
            self._add(intervalName, (start, end))

    def _add(self, name, pair):
        if self.lock is not None:
            with self.lock:
                self.pairs.setdefault(name, []).append(pair)
        else:
            self.pairs.setdefault(name, []).append(pair)

    def getDurationsForName(self, intervalName):
        if intervalName not in self.pairs:
            return None
        return [(e - s) for (s, e) in self.pairs[intervalName]]

    def merge(self, other):
        for (name, lst) in other.pairs.items():
            for pair in lst:
                self._add(name, pair)

    def makeSummaryDict(self):
        d = {}
        for name in self.pairs:
            dur = numpy.array(self.getDurationsForName(name))
            d[name] = {'total': float(dur.sum()), 'min': float(dur.min()), 'max': float(dur.max()),
                'mean': float(dur.mean()), 'count': int(len(dur))}
            # the fraction of the elapsed wall time in which at least one such interval was open
            events = sorted([(s, 1) for (s, e) in self.pairs[name]] + [(e, -1) for (s, e) in self.pairs[name]])
            (open_, last, busy) = (0, None, 0.0)
            for (t, delta) in events:
                if open_ > 0:
                    busy += t - last
                open_ += delta
                last = t
            d[name]['busy'] = busy
        return d

    def __getstate__(self):
        return {'pairs': self.pairs, 'withLock': self.lock is not None}

    def __setstate__(self, state):
        self.__init__(pairs=state['pairs'], withLock=state['withLock'])
