"""CPU: the C-ABI library loads and exports every symbol include/apgs_b200.h declares
(no compute calls -- there is no GPU here), and argument validation happens before launch."""
import ctypes as C
import os
import re

import pytest

from apgs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    out = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if fn.endswith(".h"):
            txt = open(os.path.join(ROOT, "include", fn)).read()
            out += re.findall(r"APGS_API\s+[\w\s\*]+?\b(apgs_\w+)\s*\(", txt)
    return out


def test_header_symbols_exported_and_bound():
    names = _declared()
    assert len(names) >= 9
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), "symbol %s missing from %s" % (n, _lib.LIB_PATH)
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    for n in _lib.SIGNATURES:
        assert n in names, "%s bound but not declared in include/*.h" % n


def test_version_and_pure_host_queries():
    lib = _lib.load()
    assert lib.apgs_version() == 2
    assert lib.apgs_ancestors_workspace_bytes(1000, 1024) == 0            # CDF tile lives in smem
    need = lib.apgs_ancestors_workspace_bytes(100000, 10)                 # transposed CDF + per-chunk partials
    assert 100000 * 10 * 4 <= need <= 100000 * 10 * 4 + (1 << 16) and need % 256 == 0
    assert lib.apgs_ancestors_workspace_bytes(8192, 7) == 0
    assert lib.apgs_resample_workspace_bytes(1000, 1024) >= 1000 * 1024 * 4


def test_bad_arguments_are_rejected_before_launch():
    lib = _lib.load()
    # shape / null checks return a negative code and set the thread-local message; nothing is launched
    assert lib.apgs_weight_stats_f32(None, 0, 4, None, None, None, None, 0, None) < 0
    assert b"positive" in lib.apgs_last_error()
    assert lib.apgs_weight_stats_f32(None, 4, 4, None, None, None, None, 0, None) < 0
    assert lib.apgs_sample_ancestors_f32(None, None, 4, 4, 0, None, 0, 0, 0, None, None, None, None, None, None, 0, None) < 0
    assert lib.apgs_sample_ancestors_f32(C.c_void_p(16), None, 4, 4, 7, C.c_void_p(16), 0, 0, 0, C.c_void_p(16), None,
                                         None, None, None, None, 0, None) < 0
    assert b"strategy" in lib.apgs_last_error()
    assert lib.apgs_gather_ancestors(None, None, 4, 4, None, 0, None) < 0
    assert lib.apgs_resample_f32(C.c_void_p(16), None, 4, 4, 0, C.c_void_p(16), 0, 0, 0, None, 0, None, None, None, None,
                                 None, 0, None) == -4     # APGS_E_WORKSPACE


def test_no_cpu_fallback():
    import torch
    from apgs_b200 import ops
    from apgs_b200.resampler import Resampler
    with pytest.raises(_lib.ApgsError):
        ops.weight_stats(torch.zeros(4, 4))
    with pytest.raises(_lib.ApgsError):
        Resampler("systematic", 4, False, None)
