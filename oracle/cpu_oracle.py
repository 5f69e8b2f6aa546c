"""Python face of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  ``dfgnn_b200/`` never does.

Two independent restatements live here:

* the C library ``dfgnn_oracle.c`` (fp32 and fp64 instantiations, OpenMP) whose
  loops follow the fused reference kernels line by line, and
* ``*_coo`` functions in plain numpy that follow the reference's *non-fused*
  DGL-sparse formulation (``bsddmm -> softmax -> bspmm`` on an unsorted COO,
  ``DFGNN/layers/GT/gtconv_layer.py:29-33``, ``GAT/gatconv_layer.py:30-38``,
  ``AGNN/agnn_layer.py:14-19``) with scatter ops, so that the two can be checked
  against each other without sharing code.

Parity status: pinned against outputs of the reference's own CUDA kernels
(``oracle/_ref``) on B200, committed as ``tests/golden/*.npz`` together with the
generating script ``tests/golden/make_golden.py``; index formats are
additionally pinned against ``scipy.sparse``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_lib = None

_I32P = ctypes.POINTER(ctypes.c_int32)
_I64P = ctypes.POINTER(ctypes.c_int64)


def build() -> None:
    """Compile the C oracle (gcc, a few seconds)."""
    subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def _cpu_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return " avx2" in line
    except OSError:
        pass
    return False


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        name = "libdfgnn_oracle_avx2.so" if _cpu_has_avx2() else "libdfgnn_oracle.so"
        path = os.path.join(_BUILD, name)
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _real(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError(dtype)


def _c(a, dtype) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=dtype)


# ----------------------------------------------------------------------------- #
# index formats                                                                 #
# ----------------------------------------------------------------------------- #

def coo_to_csr(row, col, n: int):
    """-> row_ptr[n+1], col_ind[E], rows[E], perm[E]  (all int32). See dfgnn_oracle.c."""
    row = _c(row, np.int64)
    col = _c(col, np.int64)
    nnz = row.shape[0]
    row_ptr = np.empty(n + 1, np.int32)
    col_ind = np.empty(nnz, np.int32)
    rows = np.empty(nnz, np.int32)
    perm = np.empty(nnz, np.int32)
    rc = lib().oracle_coo_to_csr(ctypes.c_int64(n), ctypes.c_int64(nnz), _p(row), _p(col),
                                 _p(row_ptr), _p(col_ind), _p(rows), _p(perm))
    if rc != 0:
        raise ValueError(f"oracle_coo_to_csr failed rc={rc}")
    return row_ptr, col_ind, rows, perm


def csr_to_csc(row_ptr, col_ind, n: int):
    """-> col_ptr[n+1], row_ind[E], val_idx[E]  (all int32)."""
    row_ptr = _c(row_ptr, np.int32)
    col_ind = _c(col_ind, np.int32)
    nnz = col_ind.shape[0]
    col_ptr = np.empty(n + 1, np.int32)
    row_ind = np.empty(nnz, np.int32)
    val_idx = np.empty(nnz, np.int32)
    rc = lib().oracle_csr_to_csc(ctypes.c_int64(n), ctypes.c_int64(nnz), _p(row_ptr), _p(col_ind),
                                 _p(col_ptr), _p(row_ind), _p(val_idx))
    if rc != 0:
        raise ValueError(f"oracle_csr_to_csc failed rc={rc}")
    return col_ptr, row_ind, val_idx


def smem_consume(fmt: str, max_neigh: int = 128, warp: int = 32) -> int:
    """DFGNN/layers/util.py:71 (CSR/softmax: 128) and :88 (hyper: 1024)."""
    mult = 8 if fmt.startswith("hyper") else 1
    return (max_neigh * mult + warp - 1) // warp * warp


# ----------------------------------------------------------------------------- #
# fused-kernel-order restatements (C)                                           #
# ----------------------------------------------------------------------------- #

def gt_forward(row_ptr, col_ind, val, Q, K, V, dtype=np.float32, want_attn: bool = True):
    """-> out[N,h,f], attn_edge[h,E] (or None)."""
    sfx, _ = _real(dtype)
    row_ptr = _c(row_ptr, np.int32)
    col_ind = _c(col_ind, np.int32)
    Q, K, V = _c(Q, dtype), _c(K, dtype), _c(V, dtype)
    val = None if val is None else _c(val, dtype)
    m, h, f = Q.shape
    nnz = col_ind.shape[0]
    out = np.empty((m, h, f), dtype)
    attn = np.empty((h, nnz), dtype) if want_attn else None
    getattr(lib(), "oracle_gt_forward" + sfx)(
        m, nnz, h, f, _p(row_ptr), _p(col_ind), _p(val), _p(Q), _p(K), _p(V), _p(out), _p(attn))
    return out, attn


def gt_backward(row_ptr, col_ind, col_ptr, row_ind, val_idx, Q, K, V, attn_edge, dO,
                dtype=np.float32):
    """-> dQ, dK, dV [N,h,f], grad_edge[h,E]."""
    sfx, _ = _real(dtype)
    row_ptr, col_ind = _c(row_ptr, np.int32), _c(col_ind, np.int32)
    col_ptr, row_ind, val_idx = _c(col_ptr, np.int32), _c(row_ind, np.int32), _c(val_idx, np.int32)
    Q, K, V, dO = _c(Q, dtype), _c(K, dtype), _c(V, dtype), _c(dO, dtype)
    attn_edge = _c(attn_edge, dtype)
    m, h, f = Q.shape
    nnz = col_ind.shape[0]
    dQ, dK, dV = (np.empty((m, h, f), dtype) for _ in range(3))
    ge = np.empty((h, nnz), dtype)
    getattr(lib(), "oracle_gt_backward" + sfx)(
        m, nnz, h, f, _p(row_ptr), _p(col_ind), _p(col_ptr), _p(row_ind), _p(val_idx),
        _p(Q), _p(K), _p(V), _p(attn_edge), _p(dO), _p(dQ), _p(dK), _p(dV), _p(ge))
    return dQ, dK, dV, ge


def gat_forward(attn_row, attn_col, row_ptr, col_ind, slope: float, feat, attn_drop: float = 0.0,
                edge_mask=None, dtype=np.float32):
    """-> out[N,h,f], edge_max[N,h], edge_sum[N,h]."""
    sfx, real = _real(dtype)
    row_ptr, col_ind = _c(row_ptr, np.int32), _c(col_ind, np.int32)
    ar, ac, feat = _c(attn_row, dtype), _c(attn_col, dtype), _c(feat, dtype)
    mask = None if edge_mask is None else _c(edge_mask, dtype)
    m, h, f = feat.shape
    nnz = col_ind.shape[0]
    out = np.empty((m, h, f), dtype)
    emax = np.empty((m, h), dtype)
    esum = np.empty((m, h), dtype)
    getattr(lib(), "oracle_gat_forward" + sfx)(
        m, nnz, h, f, _p(ar), _p(ac), _p(row_ptr), _p(col_ind), real(slope), _p(feat),
        real(attn_drop), _p(mask), _p(out), _p(emax), _p(esum))
    return out, emax, esum


def gat_backward(slope: float, attn_drop: float, row_ptr, col_ind, col_ptr, row_ind, permute,
                 edge_max, edge_sum, edge_mask, feat, attn_row, attn_col, dO, dtype=np.float32):
    """-> grad_feat[N,h,f], grad_attn_row[N,h], grad_attn_col[N,h]."""
    sfx, real = _real(dtype)
    row_ptr, col_ind = _c(row_ptr, np.int32), _c(col_ind, np.int32)
    col_ptr, row_ind, permute = _c(col_ptr, np.int32), _c(row_ind, np.int32), _c(permute, np.int32)
    emax, esum = _c(edge_max, dtype), _c(edge_sum, dtype)
    mask = None if edge_mask is None else _c(edge_mask, dtype)
    feat, ar, ac, dO = _c(feat, dtype), _c(attn_row, dtype), _c(attn_col, dtype), _c(dO, dtype)
    m, h, f = feat.shape
    nnz = col_ind.shape[0]
    gf = np.empty((m, h, f), dtype)
    gr = np.empty((m, h), dtype)
    gc = np.empty((m, h), dtype)
    getattr(lib(), "oracle_gat_backward" + sfx)(
        m, nnz, h, f, real(slope), real(attn_drop), _p(row_ptr), _p(col_ind), _p(col_ptr),
        _p(row_ind), _p(permute), _p(emax), _p(esum), _p(mask), _p(feat), _p(ar), _p(ac), _p(dO),
        _p(gf), _p(gr), _p(gc))
    return gf, gr, gc


def gat_attn_weight(a_l, a_r, feat, dtype=np.float32):
    sfx, _ = _real(dtype)
    feat = _c(feat, dtype)
    m, h, f = feat.shape
    a_l = _c(a_l, dtype).reshape(h, f)
    a_r = _c(a_r, dtype).reshape(h, f)
    ar = np.empty((m, h), dtype)
    ac = np.empty((m, h), dtype)
    getattr(lib(), "oracle_gat_attn_weight" + sfx)(m, h, f, _p(a_l), _p(a_r), _p(feat), _p(ar), _p(ac))
    return ar, ac


def l2_normalize(H, dtype=np.float32):
    sfx, _ = _real(dtype)
    H = _c(H, dtype)
    m, h, f = H.shape
    out = np.empty_like(H)
    getattr(lib(), "oracle_l2_normalize" + sfx)(m, h, f, _p(H), _p(out))
    return out


# ----------------------------------------------------------------------------- #
# DGL-sparse-order restatements (numpy scatter ops on an unsorted COO)          #
# ----------------------------------------------------------------------------- #

def _row_softmax_coo(score: np.ndarray, row: np.ndarray, n: int) -> np.ndarray:
    """dgl.sparse softmax over the edges sharing a row; score [E, h]."""
    mx = np.full((n, score.shape[1]), -np.inf, score.dtype)
    np.maximum.at(mx, row, score)
    ex = np.exp(score - mx[row])
    sm = np.zeros((n, score.shape[1]), score.dtype)
    np.add.at(sm, row, ex)
    return ex / sm[row]


def gt_forward_coo(row, col, n: int, Q, K, V, dtype=np.float64):
    """forward_dglsp of SparseMHA (gtconv_layer.py:29-33) on [N,h,f] operands.
    -> out[N,h,f], attn[E,h] in the order of the given COO."""
    row, col = _c(row, np.int64), _c(col, np.int64)
    Q, K, V = _c(Q, dtype), _c(K, dtype), _c(V, dtype)
    score = np.einsum("ehd,ehd->eh", Q[row], K[col])  # bsddmm
    attn = _row_softmax_coo(score, row, n)  # softmax
    out = np.zeros_like(V)
    np.add.at(out, row, attn[:, :, None] * V[col])  # bspmm
    return out, attn


def gat_forward_coo(row, col, n: int, attn_row, attn_col, feat, slope: float, dtype=np.float64):
    """forward_dglsp of GATConvDGL (gatconv_layer.py:30-38)."""
    row, col = _c(row, np.int64), _c(col, np.int64)
    ar, ac, feat = _c(attn_row, dtype), _c(attn_col, dtype), _c(feat, dtype)
    e = ar[row] + ac[col]
    e = np.where(e > 0, e, e * slope)
    attn = _row_softmax_coo(e, row, n)
    out = np.zeros_like(feat)
    np.add.at(out, row, attn[:, :, None] * feat[col])
    return out, attn


def agnn_forward_coo(row, col, n: int, H, dtype=np.float64):
    """forward_dglsp of AGNNConvDGL (agnn_layer.py:14-19): Q = K = normalize(H), V = H."""
    H = _c(H, dtype)
    Hn = H / np.maximum(np.linalg.norm(H, axis=-1, keepdims=True), 1e-12)
    return gt_forward_coo(row, col, n, Hn, Hn, H, dtype)


def gt_backward_coo(row, col, n: int, Q, K, V, dO, dtype=np.float64):
    """Analytic gradient of gt_forward_coo (what autograd gives the non-fused path)."""
    row, col = _c(row, np.int64), _c(col, np.int64)
    Q, K, V, dO = _c(Q, dtype), _c(K, dtype), _c(V, dtype), _c(dO, dtype)
    _, p = gt_forward_coo(row, col, n, Q, K, V, dtype)
    dA = np.einsum("ehd,ehd->eh", dO[row], V[col])
    t = dA * p
    s = np.zeros((n, p.shape[1]), dtype)
    np.add.at(s, row, t)
    dS = t - s[row] * p
    dQ, dK, dV = np.zeros_like(Q), np.zeros_like(K), np.zeros_like(V)
    np.add.at(dQ, row, dS[:, :, None] * K[col])
    np.add.at(dK, col, dS[:, :, None] * Q[row])
    np.add.at(dV, col, p[:, :, None] * dO[row])
    return dQ, dK, dV


def gat_backward_coo(row, col, n: int, attn_row, attn_col, feat, dO, slope: float, dtype=np.float64):
    row, col = _c(row, np.int64), _c(col, np.int64)
    ar, ac, feat, dO = _c(attn_row, dtype), _c(attn_col, dtype), _c(feat, dtype), _c(dO, dtype)
    e = ar[row] + ac[col]
    _, p = gat_forward_coo(row, col, n, ar, ac, feat, slope, dtype)
    g = np.einsum("ehd,ehd->eh", dO[row], feat[col])
    w = np.zeros((n, p.shape[1]), dtype)
    np.add.at(w, row, p * g)
    de = p * (g - w[row]) * np.where(e < 0, slope, 1.0)
    gr, gc, gf = np.zeros_like(ar), np.zeros_like(ac), np.zeros_like(feat)
    np.add.at(gr, row, de)
    np.add.at(gc, col, de)
    np.add.at(gf, col, p[:, :, None] * dO[row])
    return gf, gr, gc
