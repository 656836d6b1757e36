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


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof and the offset of every field of the three structures that cross the boundary: the
    header compiled by gcc (plain C, which also shows the header is C, not C++) against the ctypes
    declarations in pyshepseg_b200/_lib.py"""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('gcc is not installed')
    structs = (('ssg_tile_params', _lib.TileParams), ('ssg_tile_result', _lib.TileResult),
        ('ssg_tile_tables', _lib.TileTables))
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "shepseg_b200.h"', 'int main(void) {']
    for (cname, cls) in structs:
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for (field, _) in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, field, cname, field))
    lines += ['printf("abi %d\\n", SSG_ABI_VERSION);', 'return 0; }']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = str(tmp_path / 'layout')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', exe],
        check=True)
    got = dict(l.rsplit(' ', 1) for l in subprocess.run([exe], check=True, stdout=subprocess.PIPE
        ).stdout.decode().splitlines())
    for (cname, cls) in structs:
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for (field, _) in cls._fields_:
            assert int(got['%s.%s' % (cname, field)]) == getattr(cls, field).offset, (cname, field)
    assert int(got['abi']) == _lib.ABI_VERSION
