"""The C-ABI library loads on a CPU-only box and exports every symbol include/sic.h declares (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from domain_specific_image_compression_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sic.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sic_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sic.h but not exported by libsic.so"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)
    assert lib.sic_version() == 100


def test_workspace_size_queries_are_pure_host_calls():
    lib = _lib.load()
    assert lib.sic_bottleneck_workspace_bytes(16, 192, 256) >= 256 + 3 * 4 * 16 * 192 * 2
    assert lib.sic_gdn_bwd_workspace_bytes(16, 128, 65536) >= 2 * 128 * 16 * 8 * 4
    assert lib.sic_bottleneck_workspace_bytes(0, 1, 1) == 256


def test_argument_errors_are_reported_not_thrown():
    lib = _lib.load()
    rc = lib.sic_bottleneck_fwd(None, None, None, None, None, None, 0, 1, 1, 0, 0, 0, None, None, None, None, 0, None)
    assert rc == -1 and b"empty shape" in lib.sic_last_error()
    rc = lib.sic_gdn_fwd(None, None, None, None, 1, 1, 1, 0, 0, None, None)
    assert rc == -1 and b"null" in lib.sic_last_error()
    rc = lib.sic_build_cdf_tables(7, None, None, 1, 1, 1, None, None, 4, None, None)
    assert rc == -1
    with pytest.raises(_lib.SicError):
        _lib.check(rc, "sic_build_cdf_tables")


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing elsewhere."""
    import torch
    import domain_specific_image_compression_b200 as sic
    from domain_specific_image_compression_b200 import functional as F
    with pytest.raises(sic.SicError):
        F.gdn(torch.randn(1, 4, 4, 4), torch.ones(4), torch.ones(4, 1, 1, 1))
    with pytest.raises(sic.SicError):
        sic.StudentT().neg_log2_prob(torch.randn(1, 4, 4, 4), torch.ones(1, 4, 1, 1), torch.full((1, 4, 1, 1), 3.0))
    with pytest.raises(sic.SicError):
        sic.CompressionModel(N=8, M=8)(torch.rand(1, 3, 32, 32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "domain_specific_image_compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
