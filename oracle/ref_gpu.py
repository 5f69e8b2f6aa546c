"""Loader for the reference's own CUDA extensions compiled for sm_100a
(``oracle/_ref``, built by ``oracle/build_ref.py``).  TEST INFRASTRUCTURE ONLY.

``fused_gtconv()`` / ``fused_gatconv()`` return the pybind modules exactly as the
reference's ``DFGNN/operators/*.py`` import them; ``None`` when they were not
built (the caller then skips the reference-kernel comparison and says so).

The reference kernels are only defined inside an envelope (SURVEY.md 8a notes):
h == 1; ``smem_consume`` >= edges per 8-row block (hyper) / >= max degree
(softmax); f % 32 == 0 for tiling.  ``hyper_smem`` / ``softmax_smem`` compute a
safe ``smem_consume`` for a given CSR.
"""
from __future__ import annotations

import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def _load(name: str):
    if name in _cache:
        return _cache[name]
    path = os.path.join(_HERE, "_ref", name, name + ".so")
    mod = None
    if os.path.exists(path):
        import torch  # noqa: F401  (the extension links against libtorch)
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules.setdefault(name, mod)
    _cache[name] = mod
    return mod


def fused_gtconv():
    return _load("fused_gtconv")


def fused_gatconv():
    return _load("fused_gatconv")


def available() -> bool:
    return fused_gtconv() is not None and fused_gatconv() is not None


def _round32(x: int) -> int:
    return max(32, (int(x) + 31) // 32 * 32)


def hyper_smem(row_ptr) -> int:
    """smallest safe smem_consume for the hyper kernels: max edges of an 8-row block."""
    import torch
    rp = row_ptr.long().cpu()
    m = rp.numel() - 1
    idx = torch.arange(0, m + 8, 8).clamp(max=m)
    blk = rp[idx[1:]] - rp[idx[:-1]]
    return _round32(int(blk.max()) if blk.numel() else 0)


def softmax_smem(row_ptr) -> int:
    import torch
    rp = row_ptr.long().cpu()
    deg = rp[1:] - rp[:-1]
    return _round32(int(deg.max()) if deg.numel() else 0)
