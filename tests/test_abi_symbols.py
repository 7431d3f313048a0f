"""CPU-only: libscz.so loads and exports every symbol include/scz.h declares; no compute calls."""
import ctypes as C
import os

import pytest

import scz_b200 as scz
from scz_b200 import binding


def test_library_exports_every_declared_symbol():
    assert os.path.exists(binding.LIB_PATH), "libscz.so not built: run __graft_entry__.build()"
    L = binding.lib()
    names = binding.declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/scz.h but not exported: {missing}"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(scz.SczError):
        scz.Context()
    # the raw C entry point refuses as well
    h = C.c_void_p()
    assert binding.lib().scz_ctx_create(0, 0, 8, None, C.byref(h)) == -5


def test_product_never_touches_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "scalable-collaborative-zksnark_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
