"""ctypes binding of the C-ABI library ``libdfgnn_b200.so`` (include/dfgnn_b200.h).

This module is the only place that touches the native library.  It fails
loudly: if the shared object is missing or a symbol is absent the import of any
operator raises -- there is no Python / torch fallback for the conv.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# DFGNN_B200_LIB selects another build of the same ABI (developer variants, tools/kbench.py)
LIB_PATH = os.environ.get("DFGNN_B200_LIB") or os.path.join(_HERE, "libdfgnn_b200.so")
CSRC = os.path.join(_HERE, "csrc")

_lib = None


class DFGNNError(RuntimeError):
    """A non-zero return code from the native library (the reference raises
    RuntimeError from TORCH_CHECK, fused_gtconv.cpp:7-13)."""


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    jobs = str(os.cpu_count() or 4)
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-C", CSRC, "-j", jobs], check=True, stdout=out)
    return LIB_PATH


_P = c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "dfgnn_abi_version": (c_int, []),
    "dfgnn_last_error": (ctypes.c_char_p, []),
    "dfgnn_launch_count": (c_uint64, []),
    "dfgnn_last_kernel": (ctypes.c_char_p, [c_int]),
    "dfgnn_format_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "dfgnn_coo_to_csr": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "dfgnn_csr_to_csc": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "dfgnn_gt_hyper_forward": (c_int, [c_int] * 4 + [_P] * 7 + [c_int] + [_P] * 5 + [_P]),
    "dfgnn_gt_backward": (c_int, [c_int] * 5 + [_P] * 7 + [c_int] + [_P] * 9 + [_P]),
    "dfgnn_gt_backward_phase": (c_int, [c_int] * 6 + [_P] * 7 + [c_int] + [_P] * 9 + [_P]),
    "dfgnn_block_plan_check": (c_int, [c_int] * 3 + [_P] * 6 + [_P]),
    "dfgnn_gt_block_supported": (c_int, [c_int] * 5),
    "dfgnn_set_block_mode": (c_int, [c_int]),
    "dfgnn_gt_dense_supported": (c_int, [c_int] * 3),
    "dfgnn_gt_dense_forward": (c_int, [c_int, _P] + [c_int] * 5 + [_P] * 7 + [_P]),
    "dfgnn_gt_dense_tc_supported": (c_int, [c_int] * 3),
    "dfgnn_tc_poison_tmem": (c_int, [_P]),
    "dfgnn_tc_balanced_lists": (c_int, [c_int, _P, c_int, c_int, _P, _P]),
    "dfgnn_block_adj_bits": (c_int, [c_int] * 4 + [_P] * 4 + [_P]),
    "dfgnn_gt_dense_tc_forward": (c_int, [c_int, _P] + [c_int] * 5 + [_P, _P, c_int, _P, _P] + [_P] * 5 + [_P]),
    "dfgnn_gt_dense_tc_backward_col": (c_int, [c_int, _P] + [c_int] * 5 + [_P, _P, c_int, _P, _P] + [_P] * 5 + [_P]),
    "dfgnn_gt_dense_tc_backward_ws_floats": (c_size_t, [c_int, c_int]),
    "dfgnn_gt_dense_tc_backward": (c_int, [c_int, c_int, _P] + [c_int] * 5 + [_P, _P, c_int, _P, _P, c_int, _P, _P] + [_P] * 10 + [_P]),
    "dfgnn_gt_block_forward": (c_int, [c_int, _P] + [c_int] * 5 + [_P] * 8 + [_P]),
    "dfgnn_gt_block_backward": (c_int, [c_int, c_int, _P] + [c_int] * 5 + [_P] * 15 + [_P]),
    "dfgnn_proj_weight_image_floats": (c_size_t, [c_int, c_int]),
    "dfgnn_proj_pack_weights": (c_int, [c_int, c_int, _P, _P, _P]),
    "dfgnn_proj_forward": (c_int, [c_int] * 4 + [_P] * 8 + [c_int] + [_P] * 4 + [_P]),
    "dfgnn_gt_backward_cols": (c_int, [c_int] * 8 + [_P] * 7 + [c_int] + [_P] * 9 + [_P]),
    "dfgnn_gt_hyper_inference": (c_int, [c_int] * 4 + [_P] * 4 + [c_int] + [_P] * 4 + [_P]),
    "dfgnn_gt_softmax_inference": (c_int, [c_int] * 4 + [_P] * 4 + [c_int] + [_P] * 4 + [_P]),
    "dfgnn_gt_softmax_gm_inference": (c_int, [c_int] * 4 + [_P] * 4 + [_P] * 4 + [_P]),
    "dfgnn_gt_tiling_inference": (c_int, [c_int] * 4 + [_P] * 3 + [c_int] + [_P] * 4 + [_P]),
    "dfgnn_gt_csr_inference": (c_int, [c_int] * 4 + [_P] * 3 + [c_int] + [_P] * 4 + [_P]),
    "dfgnn_gt_csr_gm_inference": (c_int, [c_int] * 4 + [_P] * 3 + [_P] * 4 + [_P]),
    "dfgnn_agnn_forward": (c_int, [c_int] * 4 + [_P] * 6 + [_P]),
    "dfgnn_gat_forward": (c_int, [c_int] * 4 + [_P] * 4 + [c_float, _P, c_float, c_uint64] + [_P] * 4 + [_P]),
    "dfgnn_gat_backward": (c_int, [c_int] * 5 + [c_float, c_float] + [_P] * 16 + [_P]),
    "dfgnn_gat_backward_phase": (c_int, [c_int] * 6 + [c_float, c_float] + [_P] * 16 + [_P]),
    "dfgnn_gat_backward_cols": (c_int, [c_int] * 8 + [c_float, c_float] + [_P] * 16 + [_P]),
    "dfgnn_gat_inference": (c_int, [c_int] * 4 + [_P] * 4 + [c_float, _P, _P, _P]),
    "dfgnn_gat_inference_hyper": (c_int, [c_int] * 5 + [_P] * 5 + [c_float, _P, _P, _P]),
    "dfgnn_gat_inference_hyper_recompute": (c_int, [c_int] * 4 + [_P] * 4 + [c_float, _P, _P, _P]),
    "dfgnn_gat_inference_softmax": (c_int, [c_int] * 5 + [_P] * 5 + [c_float, _P, _P, _P]),
    "dfgnn_gat_inference_softmax_gm": (c_int, [c_int] * 4 + [_P] * 5 + [c_float, _P, _P, _P]),
    "dfgnn_gat_inference_tiling": (c_int, [c_int] * 4 + [_P] * 4 + [c_float, _P, _P, _P]),
    "dfgnn_gat_inference_hyper_v2": (c_int, [c_int] * 5 + [_P] * 4 + [c_float] + [_P] * 4 + [_P]),
    "dfgnn_gat_attn_weight": (c_int, [c_int] * 3 + [_P] * 5 + [_P]),
}

EXPORTS = tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DFGNNError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C dfgnn_b200/csrc`). There is no fallback implementation.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        if handle.dfgnn_abi_version() != 1:
            raise DFGNNError("libdfgnn_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().dfgnn_last_error().decode("utf-8", "replace")
        raise DFGNNError(f"{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().dfgnn_launch_count())


def last_kernel(slot: int) -> str:
    """Kernel the last forward (0) / backward row-side (1) / column-side (2) call dispatched to."""
    return (lib().dfgnn_last_kernel(int(slot)) or b"").decode()


def source_sha() -> str:
    """sha256 over the kernel sources (csrc/*.cu, *.cuh, *.h): ties a committed ncu capture
    (profiles/*_traffic.json) to the kernels it was taken from."""
    import hashlib
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as fh:
                h.update(name.encode() + b"\0" + fh.read())
    return h.hexdigest()[:16]
