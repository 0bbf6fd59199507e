"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/cvb200.h declares; the python mirror keeps the reference's error behaviour.
No compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "cvb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cvb200_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(lib_built):
    syms = _declared_symbols()
    assert len(syms) >= 9
    L = ctypes.CDLL(lib_built)
    for s in syms:
        assert hasattr(L, s), "libcvb200.so does not export %s" % s


def test_python_binding_covers_header(lib_built):
    from canonicalvoting_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    L = _lib.load()
    assert L.cvb200_abi_version() == _lib.ABI_VERSION
    assert L.cvb200_hv_grid_dims_work_bytes() >= 128
    dims = (ctypes.c_int32 * 3)(128, 128, 128)
    wb = L.cvb200_hv_forward_work_bytes(dims)
    assert wb == 128 ** 3 * 32


def test_argument_errors_do_not_need_a_gpu(lib_built):
    from canonicalvoting_b200 import _lib
    L = _lib.load()
    corner = _lib.f3((0, 0, 0))
    rc = L.cvb200_hv_forward(None, None, None, None, 0, 0.03, 12, corner, _lib.i3((0, 4, 4)),
                             None, None, None, None, 0, None)
    assert rc == -1 and b"dims" in L.cvb200_last_error()
    rc = L.cvb200_hv_forward(None, None, None, None, 0, 0.03, 0, corner, _lib.i3((4, 4, 4)),
                             None, None, None, None, 0, None)
    assert rc == -1


def test_reference_error_behaviour_on_cpu_tensors():
    # hv_cuda.cpp:26-28: TORCH_CHECK(is_cuda) -> RuntimeError "<name> must be a CUDA tensor"
    import hv_cuda
    p = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="points must be a CUDA tensor"):
        hv_cuda.forward(p, p, p, torch.zeros(4), torch.tensor(0.03), torch.tensor(12, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="grad_grid must be a CUDA tensor"):
        hv_cuda.backward(torch.zeros(2, 2, 2), p, p, p, torch.zeros(4), torch.tensor(0.03),
                         torch.tensor(12, dtype=torch.int32))


def test_module_names_of_the_reference():
    import hough_voting
    import hv_cuda
    assert callable(hv_cuda.forward) and callable(hv_cuda.backward)
    assert callable(hough_voting.vote)
    assert hough_voting.HVFunction.forward and hough_voting.HoughVoting


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under the product packages may reference it
    bad = []
    for base in ("canonicalvoting_b200", "hv_cuda.py", "hough_voting.py", "MinkowskiEngine"):
        p = os.path.join(ROOT, base)
        files = [p] if os.path.isfile(p) else [os.path.join(d, f) for d, _, fs in os.walk(p) for f in fs
                                                if f.endswith((".py", ".cu", ".cuh", ".h"))]
        for f in files:
            if not os.path.exists(f):
                continue
            src = open(f).read()
            if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "hv_oracle" in src.replace(
                    "oracle/hv_oracle.c header", ""):
                bad.append(f)
    assert not bad, bad
