"""Python mirror of the reference's pybind modules ``fused_gtconv`` / ``fused_gatconv``.

Every function here has the name, argument order and return shape of the
corresponding ``m.def`` in the reference
(``DFGNN/src/fused_gtconv/fused_gtconv.cpp:577-602``,
``DFGNN/src/fused_gatconv/fused_gatconv.cpp:355-372``) and forwards to the
C-ABI library through ctypes with raw device pointers on torch's current stream.
Argument checking mirrors the reference's ``CHECK_DEVICE`` / ``CHECK_CONTIGUOUS``
(``fused_gtconv.cpp:7-13``) and additionally enforces the dtypes and shapes the
reference only ``assert``s (compiled out under ``-DNDEBUG``).
"""
from __future__ import annotations

from typing import List

import torch

from .. import _lib


def _chk(name: str, t: torch.Tensor, dtype=None, allow_none: bool = False):
    if t is None:
        if allow_none:
            return
        raise RuntimeError(f"{name} must not be None")
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if t.device.type != "cuda":
        raise RuntimeError(f"{name} must be on CUDA")  # CHECK_DEVICE
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")  # CHECK_CONTIGUOUS
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")


def _ptr(t):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _val_ptr(val):
    """Edge weights of the GT scores.  The preprocessing (formats.coo_to_csr) marks its all-ones
    `val` tensor; for those the kernels skip the weight loads (NULL == all ones)."""
    return None if (val is None or getattr(val, "_dfgnn_ones", False)) else _ptr(val)


def _blocks(row_ptr, m, nnz, h, f, val=None, backward=False, training=False):
    """-> (plan, algo): the block plan the preprocessing attached to row_ptr
    (formats.attach_block_plan) and the kernel family for this call -- 0 general, 1 shared-memory
    staged (block_gt.cuh), 2 dense mma.sync (dense_gt.cuh, forward only), 3 dense tcgen05
    (dense_tc.cu)."""
    from ..formats import find_block_plan
    plan = find_block_plan(row_ptr)
    if plan is None or plan.blk_ptr.device != row_ptr.device:
        return None, 0
    unweighted = val is None or getattr(val, "_dfgnn_ones", False)
    algo = plan.algorithm(m, nnz, h, f, unweighted, training or backward)
    if backward and algo == 2:  # the mma.sync dense kernels are forward kernels
        algo = 1 if plan.supported(m, nnz, h, f) else 0
    return (plan, algo) if algo else (None, 0)


def _stream(ref: torch.Tensor):
    return torch.cuda.current_stream(ref.device).cuda_stream


def _csr_dims(indptr, indices, X, fn, row_side=True):
    if X.dim() != 3:
        raise RuntimeError(f"{fn}: features must be [N, heads, dim], got {tuple(X.shape)}")
    m = indptr.numel() - 1
    if row_side and X.shape[0] != m:
        raise RuntimeError(f"{fn}: {X.shape[0]} feature rows for {m} graph rows")
    return m, indices.numel(), X.shape[1], X.shape[2]


def _check_gt(fn, indptr, indices, Q, K, V, rows=None, val=None):
    _chk("indptr", indptr, torch.int32)
    _chk("indices", indices, torch.int32)
    _chk("rows", rows, torch.int32, allow_none=True)
    _chk("val", val, torch.float32, allow_none=True)
    for n, t in (("Q", Q), ("K", K), ("V", V)):
        _chk(n, t, torch.float32)
    # square adjacency: identical shapes; a row-partitioned shard has fewer Q rows than K/V rows
    if K.shape != V.shape or Q.dim() != 3 or Q.shape[1:] != K.shape[1:]:
        raise RuntimeError(f"{fn}: Q, K, V must have the same shape")
    if val is not None and val.numel() != indices.numel():
        raise RuntimeError(f"{fn}: val and indices differ in length")
    return _csr_dims(indptr, indices, Q, fn)


# ----------------------------------------------------------------------------- #
# module `fused_gtconv`                                                         #
# ----------------------------------------------------------------------------- #

def _block_forward(plan, algo, m, nnz, h, f, row_ptr, col_ind, val, Q, K, V, out, attn):
    L = _lib.lib()
    if algo == 3:
        return L.dfgnn_gt_dense_tc_forward(plan.n_blocks, _ptr(plan.blk_ptr), plan.max_nodes, m, nnz, h, f,
                                           _ptr(row_ptr), _ptr(plan.adj_bits), plan.n_ctas, _ptr(plan.sched_ptr),
                                           _ptr(plan.sched_idx), _ptr(Q), _ptr(K), _ptr(V), _ptr(out),
                                           _ptr(attn), _stream(Q))
    if algo == 2:
        return L.dfgnn_gt_dense_forward(plan.n_blocks, _ptr(plan.blk_ptr), plan.max_nodes, m, nnz, h, f,
                                        _ptr(row_ptr), _ptr(col_ind), _ptr(Q), _ptr(K), _ptr(V), _ptr(out),
                                        _ptr(attn), _stream(Q))
    return L.dfgnn_gt_block_forward(plan.n_blocks, _ptr(plan.blk_ptr), plan.max_nodes, m, nnz, h, f,
                                    _ptr(row_ptr), _ptr(col_ind), _val_ptr(val), _ptr(Q), _ptr(K), _ptr(V),
                                    _ptr(out), _ptr(attn), _stream(Q))


def gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem_consume,
                     Q, K, V) -> List[torch.Tensor]:
    """fused_gtconv.cpp:79-116 -> [out_feat (m,h,f), attn_edge (h,nnz)]."""
    fn = "gt_hyper_forward"
    m, nnz, h, f = _check_gt(fn, row_ptr, col_ind, Q, K, V, rows, val)
    plan, algo = _blocks(row_ptr, m, nnz, h, f, val, training=True) if K.shape[0] == m else (None, 0)
    with torch.cuda.device(Q.device):
        out = torch.empty_like(Q)
        attn = torch.empty((h, nnz), dtype=torch.float32, device=Q.device)
        if plan is not None:
            rc = _block_forward(plan, algo, m, nnz, h, f, row_ptr, col_ind, val, Q, K, V, out, attn)
            _lib.check(rc, fn)
            return [out, attn]
        rc = _lib.lib().dfgnn_gt_hyper_forward(
            m, nnz, h, f, _ptr(row_ptr), _ptr(col_ind), _ptr(rows), _val_ptr(val), _ptr(col_ptr),
            _ptr(row_ind), _ptr(val_idx), int(smem_consume), _ptr(Q), _ptr(K), _ptr(V),
            _ptr(out), _ptr(attn), _stream(Q))
    _lib.check(rc, fn)
    return [out, attn]


def gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem_consume,
                Q, K, V, attn_edge, grad, *, _phases: int = 3, _buffers=None,
                _cols=None) -> List[torch.Tensor]:
    """fused_gtconv.cpp:125-172 -> [grad_Q, grad_K, grad_V].
    Keyword-only extras (no reference counterpart, used for per-kernel timing and by the
    multi-GPU operator): ``_phases`` 1 = row-side kernel only, 2 = column-side only, 3 = both;
    ``_buffers`` = (gq, gk, gv, scratch) from an earlier call, reused instead of allocating;
    ``_cols`` = (first column, number of columns[, entries]) restricts a column-side call to a
    column range (needs ``_buffers``)."""
    fn = "gt_backward"
    m, nnz, h, f = _check_gt(fn, row_ptr, col_ind, Q, K, V, rows, val)
    _chk("col_ptr", col_ptr, torch.int32)
    _chk("row_ind", row_ind, torch.int32)
    _chk("val_idx", val_idx, torch.int32)
    _chk("attn_edge", attn_edge, torch.float32)
    _chk("grad", grad, torch.float32)
    if grad.shape != Q.shape:
        raise RuntimeError(f"{fn}: grad must have the shape of Q")
    if attn_edge.numel() != h * nnz or row_ind.numel() != nnz or val_idx.numel() != nnz:
        raise RuntimeError(f"{fn}: attn_edge / row_ind / val_idx do not match nnz={nnz}")
    n = col_ptr.numel() - 1
    if K.shape[0] != n:
        raise RuntimeError(f"{fn}: col_ptr describes {n} columns but K has {K.shape[0]} rows")
    with torch.cuda.device(Q.device):
        if _buffers is not None:
            gq, gk, gv, ge = _buffers
        else:
            gq, gk, gv = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
            ge = torch.empty((h, nnz, 2), dtype=torch.float32, device=Q.device)  # scratch {dS, p}
        tail = (m, n, nnz, h, f, _ptr(row_ptr), _ptr(col_ind), _ptr(rows), _val_ptr(val), _ptr(col_ptr),
                _ptr(row_ind), _ptr(val_idx), int(smem_consume), _ptr(Q), _ptr(K), _ptr(V),
                _ptr(attn_edge), _ptr(grad), _ptr(gq), _ptr(gk), _ptr(gv), _ptr(ge), _stream(Q))
        plan, _algo = _blocks(row_ptr, m, nnz, h, f, val, backward=True) if (n == m and _cols is None) else (None, 0)
        if plan is not None and _algo == 3:
            # dense batch: both sides on tcgen05 (csrc/dense_tc.cu); dense P / dS tiles in a work array kept on
            # the plan (it must survive from a row-side call to a column-side call)
            L = _lib.lib()
            need = int(L.dfgnn_gt_dense_tc_backward_ws_floats(m, plan.n_blocks))
            ws = getattr(plan, "_dense_ws", None)
            if ws is None or ws.numel() < need or ws.device != Q.device:
                ws = plan._dense_ws = torch.empty(need, dtype=torch.float32, device=Q.device)
            nc, sp, si = plan.col_sched
            rc = L.dfgnn_gt_dense_tc_backward(
                int(_phases), plan.n_blocks, _ptr(plan.blk_ptr), plan.max_nodes, m, nnz, h, f, _ptr(row_ptr),
                _ptr(plan.adj_bits), plan.n_ctas, _ptr(plan.sched_ptr), _ptr(plan.sched_idx), nc, _ptr(sp), _ptr(si),
                _ptr(Q), _ptr(K), _ptr(V), _ptr(attn_edge), _ptr(grad), _ptr(gq), _ptr(gk), _ptr(gv), _ptr(plan.tile_ptr),
                _ptr(ws), _stream(Q))
        elif plan is not None:
            rc = _lib.lib().dfgnn_gt_block_backward(
                int(_phases), plan.n_blocks, _ptr(plan.blk_ptr), plan.max_nodes, m, nnz, h, f, _ptr(row_ptr),
                _ptr(col_ind), _val_ptr(val), _ptr(col_ptr), _ptr(row_ind), _ptr(val_idx), _ptr(Q), _ptr(K),
                _ptr(V), _ptr(attn_edge), _ptr(grad), _ptr(gq), _ptr(gk), _ptr(gv), _ptr(ge), _stream(Q))
        elif _cols is not None:
            if _buffers is None or _phases != 2:
                raise RuntimeError(f"{fn}: _cols needs _phases=2 and _buffers")
            c0, nc = int(_cols[0]), int(_cols[1])
            rc = _lib.lib().dfgnn_gt_backward_cols(c0, nc, int(_cols[2]) if len(_cols) > 2 else -1, *tail)
        else:
            rc = _lib.lib().dfgnn_gt_backward_phase(int(_phases), *tail)
    _lib.check(rc, fn)
    if _buffers is None and _phases != 3:
        return [gq, gk, gv, ge]
    return [gq, gk, gv]


def _gt_inference(cname, fn, indptr, indices, rows, val, smem_consume, Q, K, V, has_rows, has_smem):
    m, nnz, h, f = _check_gt(fn, indptr, indices, Q, K, V, rows, val)
    plan, algo = _blocks(indptr, m, nnz, h, f, val) if K.shape[0] == m else (None, 0)
    with torch.cuda.device(Q.device):
        out = torch.empty_like(Q)
        if plan is not None:
            rc = _block_forward(plan, algo, m, nnz, h, f, indptr, indices, val, Q, K, V, out, None)
            _lib.check(rc, fn)
            return out
        args = [m, nnz, h, f, _ptr(indptr), _ptr(indices)]
        if has_rows:
            args.append(_ptr(rows))
        args.append(_val_ptr(val))
        if has_smem:
            args.append(int(smem_consume))
        args += [_ptr(Q), _ptr(K), _ptr(V), _ptr(out), _stream(Q)]
        rc = getattr(_lib.lib(), cname)(*args)
    _lib.check(rc, fn)
    return out


def gt_hyper_inference(indptr, indices, rows, val, smem_consume, Q, K, V):
    """fused_gtconv.cpp:278-314 -> [out_feat]."""
    return [_gt_inference("dfgnn_gt_hyper_inference", "gt_hyper_inference", indptr, indices, rows,
                          val, smem_consume, Q, K, V, True, True)]


def gt_softmax_inference(indptr, indices, rows, val, smem_consume, Q, K, V):
    """fused_gtconv.cpp:316-352 -> [out_feat]."""
    return [_gt_inference("dfgnn_gt_softmax_inference", "gt_softmax_inference", indptr, indices,
                          rows, val, smem_consume, Q, K, V, True, True)]


def gt_softmax_gm_inference(indptr, indices, rows, val, Q, K, V):
    """fused_gtconv.cpp:354-389 -> out_feat (a bare tensor, not a list)."""
    return _gt_inference("dfgnn_gt_softmax_gm_inference", "gt_softmax_gm_inference", indptr,
                         indices, rows, val, 0, Q, K, V, True, False)


def gt_tiling_inference(indptr, indices, val, smem_consume, Q, K, V):
    """fused_gtconv.cpp:244-276 -> [out_feat]."""
    return [_gt_inference("dfgnn_gt_tiling_inference", "gt_tiling_inference", indptr, indices,
                          None, val, smem_consume, Q, K, V, False, True)]


def gt_csr_inference(indptr, indices, val, smem_consume, Q, K, V):
    """fused_gtconv.cpp:174-207 -> [out_feat]."""
    return [_gt_inference("dfgnn_gt_csr_inference", "gt_csr_inference", indptr, indices, None, val,
                          smem_consume, Q, K, V, False, True)]


def gt_csr_gm_inference(indptr, indices, val, Q, K, V):
    """fused_gtconv.cpp:209-242 -> [out_feat]."""
    return [_gt_inference("dfgnn_gt_csr_gm_inference", "gt_csr_gm_inference", indptr, indices,
                          None, val, 0, Q, K, V, False, False)]


def agnn_forward(indptr, indices, H, want_attn: bool = False):
    """AGNN with F.normalize fused into the conv (no reference export; see dfgnn_b200.h).
    -> out_feat, or (out_feat, attn_edge) when want_attn."""
    fn = "agnn_forward"
    _chk("indptr", indptr, torch.int32)
    _chk("indices", indices, torch.int32)
    _chk("H", H, torch.float32)
    m, nnz, h, f = _csr_dims(indptr, indices, H, fn)
    with torch.cuda.device(H.device):
        out = torch.empty_like(H)
        rn = torch.empty((m, h), dtype=torch.float32, device=H.device)
        attn = torch.empty((h, nnz), dtype=torch.float32, device=H.device) if want_attn else None
        rc = _lib.lib().dfgnn_agnn_forward(m, nnz, h, f, _ptr(indptr), _ptr(indices), _ptr(H),
                                           _ptr(rn), _ptr(out), _ptr(attn), _stream(H))
    _lib.check(rc, fn)
    return (out, attn) if want_attn else out


# ----------------------------------------------------------------------------- #
# module `fused_gatconv`                                                        #
# ----------------------------------------------------------------------------- #

def _check_gat(fn, attn_row, attn_col, indptr, indices, in_feat, rows=None):
    _chk("attn_row", attn_row, torch.float32)
    _chk("attn_col", attn_col, torch.float32)
    _chk("indptr", indptr, torch.int32)
    _chk("indices", indices, torch.int32)
    _chk("rows", rows, torch.int32, allow_none=True)
    _chk("in_feat", in_feat, torch.float32)
    # in_feat / attn_col are column-side (n rows), attn_row / out are row-side (m rows)
    m, nnz, h, f = _csr_dims(indptr, indices, in_feat, fn, row_side=False)
    if attn_row.numel() != m * h or attn_col.numel() != in_feat.shape[0] * h:
        raise RuntimeError(f"{fn}: attn_row must be [rows, heads] and attn_col [cols, heads]")
    return m, nnz, h, f


_seed_counter = [0x5DEECE66D]


def _next_seed() -> int:
    # a fresh mask every call, reproducible under torch.manual_seed
    _seed_counter[0] = (_seed_counter[0] * 6364136223846793005 + 1442695040888963407) % (1 << 64)
    return (_seed_counter[0] ^ int(torch.initial_seed())) % (1 << 64)


def gat_forward(attn_row, attn_col, row_ptr, col_ind, negative_slope, in_feat, attn_drop,
                seed=None) -> List[torch.Tensor]:
    """fused_gatconv.cpp:11-32 -> [out_feat, edge_max, edge_sum, edge_mask]."""
    fn = "gat_forward"
    m, nnz, h, f = _check_gat(fn, attn_row, attn_col, row_ptr, col_ind, in_feat)
    dev = in_feat.device
    with torch.cuda.device(dev):
        out = torch.empty((m, h, f), dtype=torch.float32, device=dev)
        emax = torch.empty((m, h), dtype=torch.float32, device=dev)
        esum = torch.empty((m, h), dtype=torch.float32, device=dev)
        emask = torch.empty((nnz, h), dtype=torch.float32, device=dev)
        rc = _lib.lib().dfgnn_gat_forward(
            m, nnz, h, f, _ptr(attn_row), _ptr(attn_col), _ptr(row_ptr), _ptr(col_ind),
            float(negative_slope), _ptr(in_feat), float(attn_drop),
            _next_seed() if seed is None else int(seed) % (1 << 64),
            _ptr(out), _ptr(emax), _ptr(esum), _ptr(emask), _stream(in_feat))
    _lib.check(rc, fn)
    return [out, emax, esum, emask]


def gat_backward(negative_slope, attn_drop, row_ptr, col_ind, col_ptr, row_ind, permute,
                 edge_max, edge_sum, edge_mask, in_feat, attn_row, attn_col, grad, *,
                 _phases: int = 3, _buffers=None, _cols=None):
    """fused_gatconv.cpp:291-353 -> [grad_feat, grad_attn_row, grad_attn_col].
    ``_phases`` / ``_buffers`` / ``_cols``: see gt_backward (buffers = (gf, gr, gc, scratch))."""
    fn = "gat_backward"
    m, nnz, h, f = _check_gat(fn, attn_row, attn_col, row_ptr, col_ind, in_feat)
    for n, t in (("col_ptr", col_ptr), ("row_ind", row_ind), ("permute", permute)):
        _chk(n, t, torch.int32)
    for n, t in (("edge_max", edge_max), ("edge_sum", edge_sum), ("edge_mask", edge_mask),
                 ("grad", grad)):
        _chk(n, t, torch.float32)
    n = col_ptr.numel() - 1
    if tuple(grad.shape) != (m, h, f):
        raise RuntimeError(f"{fn}: grad must be [rows, heads, dim]")
    if row_ind.numel() != nnz or permute.numel() != nnz or in_feat.shape[0] != n:
        raise RuntimeError(f"{fn}: CSC arrays do not match the CSR")
    dev = in_feat.device
    with torch.cuda.device(dev):
        if _buffers is not None:
            gf, gr, gc, ge = _buffers
        else:
            gf = torch.empty_like(in_feat)
            gr = torch.empty((m, h), dtype=torch.float32, device=dev)
            gc = torch.empty((n, h), dtype=torch.float32, device=dev)
            ge = torch.empty((nnz, h, 2), dtype=torch.float32, device=dev)  # scratch {de, keep-scaled p}
        tail = (m, n, nnz, h, f, float(negative_slope), float(attn_drop), _ptr(row_ptr), _ptr(col_ind),
                _ptr(col_ptr), _ptr(row_ind), _ptr(permute), _ptr(edge_max), _ptr(edge_sum),
                _ptr(edge_mask), _ptr(in_feat), _ptr(attn_row), _ptr(attn_col), _ptr(grad), _ptr(gf),
                _ptr(gr), _ptr(gc), _ptr(ge), _stream(in_feat))
        if _cols is not None:
            if _buffers is None or _phases != 2:
                raise RuntimeError(f"{fn}: _cols needs _phases=2 and _buffers")
            c0, nc = int(_cols[0]), int(_cols[1])
            rc = _lib.lib().dfgnn_gat_backward_cols(c0, nc, int(_cols[2]) if len(_cols) > 2 else -1, *tail)
        else:
            rc = _lib.lib().dfgnn_gat_backward_phase(int(_phases), *tail)
    _lib.check(rc, fn)
    if _buffers is None and _phases != 3:
        return [gf, gr, gc, ge]
    return [gf, gr, gc]


def _gat_inference(cname, fn, smem_consume, attn_row, attn_col, indptr, indices, rows,
                   negative_slope, in_feat, has_smem, has_rows):
    m, nnz, h, f = _check_gat(fn, attn_row, attn_col, indptr, indices, in_feat, rows)
    with torch.cuda.device(in_feat.device):
        out = torch.empty((m, h, f), dtype=torch.float32, device=in_feat.device)
        args = [int(smem_consume)] if has_smem else []
        args += [m, nnz, h, f, _ptr(attn_row), _ptr(attn_col), _ptr(indptr), _ptr(indices)]
        if has_rows:
            args.append(_ptr(rows))
        args += [float(negative_slope), _ptr(in_feat), _ptr(out), _stream(in_feat)]
        rc = getattr(_lib.lib(), cname)(*args)
    _lib.check(rc, fn)
    return out


def gat_inference(attn_row, attn_col, row_ptr, col_ind, negative_slope, in_feat):
    """fused_gatconv.cpp:225-254."""
    return _gat_inference("dfgnn_gat_inference", "gat_inference", 0, attn_row, attn_col, row_ptr,
                          col_ind, None, negative_slope, in_feat, False, False)


def gat_inference_hyper(smem_consume, attn_row, attn_col, indptr, indices, rows, negative_slope,
                        in_feat):
    """fused_gatconv.cpp:99-124."""
    return _gat_inference("dfgnn_gat_inference_hyper", "gat_inference_hyper", smem_consume,
                          attn_row, attn_col, indptr, indices, rows, negative_slope, in_feat,
                          True, True)


def gat_inference_hyper_recompute(attn_row, attn_col, indptr, indices, negative_slope, in_feat):
    """fused_gatconv.cpp:126-150."""
    return _gat_inference("dfgnn_gat_inference_hyper_recompute", "gat_inference_hyper_recompute",
                          0, attn_row, attn_col, indptr, indices, None, negative_slope, in_feat,
                          False, False)


def gat_inference_softmax(smem_consume, attn_row, attn_col, indptr, indices, rows, negative_slope,
                          in_feat):
    """fused_gatconv.cpp:40-68."""
    return _gat_inference("dfgnn_gat_inference_softmax", "gat_inference_softmax", smem_consume,
                          attn_row, attn_col, indptr, indices, rows, negative_slope, in_feat,
                          True, True)


def gat_inference_softmax_gm(attn_row, attn_col, indptr, indices, rows, negative_slope, in_feat):
    """fused_gatconv.cpp:70-97."""
    return _gat_inference("dfgnn_gat_inference_softmax_gm", "gat_inference_softmax_gm", 0,
                          attn_row, attn_col, indptr, indices, rows, negative_slope, in_feat,
                          False, True)


def gat_inference_tiling(attn_row, attn_col, row_ptr, col_ind, negative_slope, in_feat):
    """fused_gatconv.cpp:196-223."""
    return _gat_inference("dfgnn_gat_inference_tiling", "gat_inference_tiling", 0, attn_row,
                          attn_col, row_ptr, col_ind, None, negative_slope, in_feat, False, False)


def gat_inference_hyper_v2(smem_consume, a_l, a_r, indptr, indices, negative_slope, in_feat):
    """fused_gatconv.cpp:152-166: attention logits are computed inside the call from
    a_l / a_r (any shape holding heads*dim floats in [head, dim] order)."""
    fn = "gat_inference_hyper_v2"
    _chk("a_l", a_l, torch.float32)
    _chk("a_r", a_r, torch.float32)
    _chk("indptr", indptr, torch.int32)
    _chk("indices", indices, torch.int32)
    _chk("in_feat", in_feat, torch.float32)
    m, nnz, h, f = _csr_dims(indptr, indices, in_feat, fn)
    if a_l.numel() != h * f or a_r.numel() != h * f:
        raise RuntimeError(f"{fn}: a_l / a_r must hold heads*dim = {h * f} floats")
    dev = in_feat.device
    with torch.cuda.device(dev):
        out = torch.empty_like(in_feat)
        ar = torch.empty((m, h), dtype=torch.float32, device=dev)
        ac = torch.empty((m, h), dtype=torch.float32, device=dev)
        rc = _lib.lib().dfgnn_gat_inference_hyper_v2(
            int(smem_consume), m, nnz, h, f, _ptr(a_l), _ptr(a_r), _ptr(indptr), _ptr(indices),
            float(negative_slope), _ptr(in_feat), _ptr(ar), _ptr(ac), _ptr(out), _stream(in_feat))
    _lib.check(rc, fn)
    return out


def gat_attn_weight(a_l, a_r, in_feat):
    """attn_row, attn_col = <a_l, feat>, <a_r, feat>  (fused_gatconv_hyper_v2.cu:212-250)."""
    fn = "gat_attn_weight"
    _chk("a_l", a_l, torch.float32)
    _chk("a_r", a_r, torch.float32)
    _chk("in_feat", in_feat, torch.float32)
    m, h, f = in_feat.shape
    if a_l.numel() != h * f or a_r.numel() != h * f:
        raise RuntimeError(f"{fn}: a_l / a_r must hold heads*dim = {h * f} floats")
    dev = in_feat.device
    with torch.cuda.device(dev):
        ar = torch.empty((m, h), dtype=torch.float32, device=dev)
        ac = torch.empty((m, h), dtype=torch.float32, device=dev)
        rc = _lib.lib().dfgnn_gat_attn_weight(m, h, f, _ptr(a_l), _ptr(a_r), _ptr(in_feat),
                                              _ptr(ar), _ptr(ac), _stream(in_feat))
    _lib.check(rc, fn)
    return ar, ac
