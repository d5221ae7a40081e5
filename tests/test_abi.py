"""The C-ABI library loads on a machine without a GPU and exports every symbol the public headers
declare; the product path fails loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", "ofdg", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ofdg_[a-z0-9_]+)\s*\(", text)))


def test_exports_match_headers(ofdg):
    L = ctypes.CDLL(ofdg.LIB_PATH)
    for header, listed in (("ofdg.h", ofdg.EXPORTS), ("layer.h", ofdg.LAYER_EXPORTS)):
        declared = _declared(header)
        assert declared, header
        for name in declared:
            assert hasattr(L, name), f"{name} declared in include/ofdg/{header} but not exported"
        assert sorted(listed) == declared, f"python binding list out of date for {header}"


def test_version_and_error_channel(ofdg):
    L = ofdg.lib()
    assert L.ofdg_version() == 100
    h = ctypes.c_void_p()
    rc = L.ofdg_params_create(99, 512, 384, 0, 0, 0, ctypes.byref(h))
    assert rc != 0 and b"BAD MODE" in L.ofdg_last_error()


def test_no_cpu_fallback(ofdg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ofdg.OfdgError, match="no CUDA device|CUDA"):
        ofdg.Generator(device=0, mode=1)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the product package may import, include or load it."""
    pkg = os.path.join(ROOT, "optical-flow-2d-data-generation_b200")
    bad = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|#include\s+\"[^\"]*oracle|oracle[/\\]binding)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".inc")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not bad.search(text), f"{f} references the oracle"
    root_shim = open(os.path.join(ROOT, "ofdg_b200.py")).read()
    assert not bad.search(root_shim)


def test_host_expand_routine_matches_numpy(ofdg):
    """The host half of the uint8 transport (csrc/host/expand.cpp) for every head/tail alignment."""
    import numpy as np
    rng = np.random.default_rng(5)
    src_all = rng.integers(0, 256, 5000, dtype=np.uint8)
    dst_all = np.empty(5000 + 64, np.float32)
    for streaming in (True, False):
        for off in range(0, 17):
            for n in (0, 1, 7, 15, 16, 17, 63, 64, 65, 1000, 4099):
                dst_all[:] = -1
                dst = dst_all[off:off + n]
                ofdg.expand_host(src_all[3:3 + n], dst, streaming)
                assert np.array_equal(dst, src_all[3:3 + n].astype(np.float32))
                assert np.all(dst_all[:off] == -1) and np.all(dst_all[off + n:] == -1)
