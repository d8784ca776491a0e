"""CPU tests: the C-ABI library loads and exports every symbol include/ttb200.h declares."""

import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ttb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ttb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ("ttb_inner_f64", "ttb_round_f64", "ttb_gemm_f64", "ttb_version", "ttb_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from tensor_networks_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    handle = _lib.lib()
    for s in declared_symbols():
        assert hasattr(handle, s), f"libttb200.so does not export {s}"
    assert b"ttb200" in handle.ttb_version()


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tensor_networks_b200 import TensorTrain

    with pytest.raises(RuntimeError):
        TensorTrain.rand([2, 2], [2])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "tensor_networks_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
