"""CPU: the C-ABI library loads and exports exactly the symbols include/mimosa_b200.h declares; without a GPU
the product fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mimosa_b200.h")).read()
    return sorted(set(re.findall(r"MB_API\s+[\w\s\*]+?\b(mb_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    from mimosa_b200 import capi

    assert declared_symbols() == sorted(capi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from mimosa_b200 import capi

    lib = capi.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.mb_version() >= 100


def test_struct_layouts_match_oracle_binding(oracle):
    from mimosa_b200 import capi

    for a, b in ((capi.IcpConfig, oracle.IcpConfig), (capi.Linearization, oracle.Linearization), (capi.IcpTrace, oracle.IcpTrace)):
        assert C.sizeof(a) == C.sizeof(b)
        assert [(n, t) for n, t in a._fields_] == [(n, t) for n, t in b._fields_]
    lib = capi.load()
    L = oracle.lib()
    for which, s in enumerate((capi.IcpConfig, capi.Linearization, capi.IcpTrace)):
        assert lib.mb_sizeof(which) == C.sizeof(s) == L.orc_sizeof(which)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mimosa_b200 import capi

    lib = capi.load()
    h = C.c_void_p()
    assert lib.mb_init(0, C.byref(h)) == capi.MB_ERR_NO_DEVICE
    assert b"no CPU path" in lib.mb_last_error()
    import mimosa_b200

    with pytest.raises(capi.MimosaError):
        mimosa_b200.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mimosa_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle_py" not in txt and "liboracle" not in txt, fn
                assert not re.search(r"#\s*include[^\n]*(oracle|_ref\.hpp)", txt), fn
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), fn


def test_bench_algorithmic_bytes_count():
    """bench.py's compulsory-byte count of a k-NN launch on a hand-made map: two occupied voxels in one 4x4x4 block, one
    query whose 19-voxel neighbourhood contains both."""
    import numpy as np

    import bench

    coords = np.array([[0, 0, 0], [1, 0, 0], [9, 9, 9]], np.int32)
    counts = np.array([3, 7, 20], np.int32)
    q = np.array([[0.5, 0.5, 0.5]])
    total, parts = bench.knn_algorithmic_bytes(q, coords, counts, 5, nbr_mode=19, leaf=1.0)
    # neighbourhood cells -1..1 straddle blocks -1 and 0 on every axis, minus the all-(-1) corner block that only
    # the excluded (-1,-1,-1) corner touches: 7 blocks
    assert parts["distinct_buckets"] == 2 and parts["bucket_bytes"] == (3 + 7) * 16 and parts["distinct_blocks"] == 7
    assert total == 24 + 7 * 32 + 2 * 4 + 160 + (5 * 16 + 1)
    assert parts["survey_8d_bytes"] == 16 + 19 * 16 + 2 * 336 + 5 * 16
