"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

from pyshepseg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'shepseg_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ssg_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_functions():
    names = declared_functions()
    assert 'ssg_segment_tile' in names and 'ssg_tile_tables_device' in names
    assert len(names) >= 30


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIBPATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, 'declared in the header but not exported: %s' % missing


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES.keys()) == declared_functions()


def test_no_cpu_fallback():
    lib = _lib.load()
    assert lib.ssg_abi_version() == 2
    if lib.ssg_device_count() == 0:
        with pytest.raises(_lib.ShepsegB200Error):
            _lib.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'pyshepseg_b200')
    for (d, _, files) in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(d, f)).read()
                assert 'oracle' not in src.replace('the oracle', '').replace('oracle /', '').replace(
                    "oracle's", ''), f
