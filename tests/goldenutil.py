"""Helpers to load the golden fixtures written by tests/golden/make_golden.py."""
import glob
import json
import os

import numpy

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


class Centres(object):
    """Minimal stand-in for a fitted sklearn KMeans: the path only needs the centres."""
    def __init__(self, centres):
        self.cluster_centers_ = numpy.ascontiguousarray(centres, dtype=numpy.float64)


def _load(path):
    z = numpy.load(path)
    d = dict((k, z[k]) for k in z.files)
    d['meta'] = json.loads(str(d['meta']))
    d['name'] = os.path.basename(path)[:-4]
    return d


def single_tile_names():
    names = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, '*.npz'))):
        n = os.path.basename(p)[:-4]
        if n.startswith(('tiled_', 'stats_')) or n == 'clump_only':
            continue
        names.append(n)
    return names


def tiled_names():
    return sorted(os.path.basename(p)[:-4]
        for p in glob.glob(os.path.join(GOLDEN_DIR, 'tiled_*.npz')))


def load(name):
    return _load(os.path.join(GOLDEN_DIR, name + '.npz'))


def msd_of(case):
    """The maxSpectralDiff argument as the reference call received it."""
    return case['meta']['maxSpectralDiff']


def resolved_msd(case):
    """The resolved maxSpectralDiff with the dtype the reference produced."""
    v = float(case['msd_value'])
    if int(case['msd_is_f32']):
        return numpy.float32(v)
    m = case['meta']['maxSpectralDiff']
    if isinstance(m, int) and not isinstance(m, bool):
        return int(m)
    return v
