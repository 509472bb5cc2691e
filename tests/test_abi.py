"""The C-ABI library loads on a CPU-only box and exports every symbol include/quandary_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

import quandary_b200 as qb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "quandary_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = qb.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/quandary_b200.h but not exported"
    # and the ctypes mirror binds all of them
    assert set(names) == set(lib._qb_symbols), set(names) ^ set(lib._qb_symbols)


def test_struct_layout_matches_header():
    """sizeof of the ctypes mirrors == sizeof in C (compiled with gcc from the header)."""
    import subprocess, tempfile
    code = '#include "quandary_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(qb_segment), sizeof(qb_system_desc), sizeof(qb_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(code)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [ctypes.sizeof(qb.Segment), ctypes.sizeof(qb.SystemDesc), ctypes.sizeof(qb.Stats)]


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device engine creation fails loudly; with one, it succeeds."""
    from qb_systems import cfg_system
    desc, _ = cfg_system("C2", ntime=4)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except ImportError:
        has_gpu = False
    if has_gpu:
        qb.Engine(desc).close()
    else:
        with pytest.raises(qb.QBError):
            qb.Engine(desc)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(qb.QBError):
        qb.load_library(str(tmp_path / "libquandary_b200.so"))
