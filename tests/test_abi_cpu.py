"""The C-ABI library loads without a GPU, exports every symbol the header
declares with the arity the ctypes table binds, and rejects bad arguments before
touching the device.  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

from dfgnn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dfgnn_b200.h")


def _header_decls():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|size_t|uint64_t|const char \*)\s*\*?(dfgnn_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_header_and_ctypes_table_agree():
    decls = _header_decls()
    assert set(decls) == set(_lib.EXPORTS), set(decls) ^ set(_lib.EXPORTS)
    for name, n in decls.items():
        assert len(_lib._SIGNATURES[name][1]) == n, f"{name}: header has {n} args"


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libdfgnn_b200.so not built (run __graft_entry__.build())"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = set(_header_decls()) - exported
    assert not missing, f"not exported: {missing}"
    # nothing else leaks (built with -fvisibility=hidden)
    assert all(s.startswith("dfgnn_") for s in exported), exported


def test_library_is_sm100a_sass():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_loads_and_reports_version():
    L = _lib.lib()
    assert L.dfgnn_abi_version() == 1
    assert _lib.launch_count() == 0 or _lib.launch_count() > 0


def test_argument_validation_happens_before_any_device_work():
    L = _lib.lib()
    n0 = _lib.launch_count()
    dummy = ctypes.c_void_p(16)
    # f > 512
    rc = L.dfgnn_gt_hyper_inference(4, 4, 1, 1024, dummy, dummy, None, None, 0, dummy, dummy, dummy, dummy, None)
    assert rc == -2 and b"not supported" in L.dfgnn_last_error()
    # NULL required pointer
    rc = L.dfgnn_gt_hyper_inference(4, 4, 1, 64, None, dummy, None, None, 0, dummy, dummy, dummy, dummy, None)
    assert rc == -1 and b"NULL" in L.dfgnn_last_error()
    # negative sizes
    rc = L.dfgnn_gat_inference(-1, 0, 1, 64, dummy, dummy, dummy, dummy, 0.2, dummy, dummy, None)
    assert rc == -1
    # dropout probability out of range
    rc = L.dfgnn_gat_forward(4, 4, 1, 64, dummy, dummy, dummy, dummy, 0.2, dummy, 1.5, 0, dummy, dummy, dummy, dummy, None)
    assert rc == -1 and b"attn_drop" in L.dfgnn_last_error()
    with pytest.raises(_lib.DFGNNError, match="attn_drop"):
        _lib.check(rc, "gat_forward")
    assert _lib.launch_count() == n0  # nothing was launched


def test_operators_refuse_cpu_tensors():
    """No CPU fallback: host tensors are an error (CHECK_DEVICE, fused_gtconv.cpp:7-8)."""
    import torch
    from dfgnn_b200.operators import GATConvFuse_inference_softmax, GTConvFuse_inference_hyper
    rp = torch.tensor([0, 1, 2], dtype=torch.int32)
    ci = torch.tensor([1, 0], dtype=torch.int32)
    x = torch.randn(2, 1, 32)
    with pytest.raises(RuntimeError, match="must be on CUDA"):
        GTConvFuse_inference_hyper(rp, ci, ci, torch.ones(2), 1024, x, x, x)
    with pytest.raises(RuntimeError, match="must be on CUDA"):
        GATConvFuse_inference_softmax(128, torch.randn(2, 1), torch.randn(2, 1), rp, ci, ci, 0.2, x)
    from dfgnn_b200 import formats
    with pytest.raises(RuntimeError, match="CUDA"):
        formats.csr_to_csc(rp, ci)
