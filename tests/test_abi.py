"""The C-ABI library loads and exports every symbol that include/temfpy_b200.h declares (no
compute calls without a GPU), and the package refuses to run without the CUDA build."""
import ctypes
import os
import re

import pytest

from temfpy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "temfpy_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmf_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_table_agree():
    assert set(header_symbols()) == set(_lib.SIGNATURES)


def test_cuda_library_exports_every_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.tmf_is_cuda() == 1
    assert lib.tmf_version() >= 100


def test_descriptor_struct_sizes():
    assert ctypes.sizeof(_lib.GemmJob) == 128
    assert ctypes.sizeof(_lib.SiteJob) == 128
    assert ctypes.sizeof(_lib.MinorBlock) == 64
    assert ctypes.sizeof(_lib.SitePlan) == 72


def test_no_cpu_fallback():
    """Without a GPU the public path must fail loudly (never route through oracle/ or the simulator)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from temfpy_b200 import slater
    import numpy as np
    with pytest.raises(RuntimeError):
        slater.C_to_MPS(np.eye(4) * 0.5, {"chi_max": 4})
    src = ""
    pkg = os.path.join(ROOT, "temfpy_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src += open(os.path.join(pkg, f)).read()
    assert "slater_oracle" not in src and "hostsim" not in src.replace("tests/hostsim", "")
