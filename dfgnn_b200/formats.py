"""Index-format construction on the GPU (COO -> CSR / CSR+COO "hyper" / CSC).

Host-side mirror of what the reference gets from ``dgl.sparse``
(``DFGNN/layers/util.py:52-162``): ``A.csr()``, ``torch.sort(A.row)``,
``A.val[val_idx]`` and ``dglsp.from_csr(...).csc()``.  The work is done by
``dfgnn_coo_to_csr`` / ``dfgnn_csr_to_csc`` of the C-ABI library (stable radix
sorts, bit-exact by definition -- SURVEY.md 8c).  There is no CPU path.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib


class SparseMatrix:
    """The few attributes of ``dgl.sparse.SparseMatrix`` the reference reads
    (``A.row``, ``A.col``, ``A.val``, ``A.shape``): the raw COO of the graph."""

    def __init__(self, row: torch.Tensor, col: torch.Tensor, shape: Tuple[int, int]):
        self.row = row
        self.col = col
        self.shape = tuple(shape)
        self.val = torch.ones(row.numel(), dtype=torch.float32, device=row.device)
        self.val._dfgnn_ones = True

    @property
    def nnz(self) -> int:
        return int(self.row.numel())

    @property
    def device(self):
        return self.row.device

    def csr(self):
        """-> (indptr, indices, value_indices) like dgl.sparse (int64, as DGL returns)."""
        row_ptr, col_ind, _, perm, _ = coo_to_csr(self.row, self.col, self.shape[0], self.shape[1])
        return row_ptr.long(), col_ind.long(), perm.long()


def _require_cuda(t: torch.Tensor, name: str):
    if t.device.type != "cuda":
        raise RuntimeError(f"{name} must be on CUDA: format construction runs on the GPU only")


def coo_to_csr(row: torch.Tensor, col: torch.Tensor, n: int, n_cols: int = None):
    """-> row_ptr[n+1] i32, col_ind[E] i32, rows[E] i32, perm[E] i32, val[E] f32 (ones).
    n rows; n_cols columns (default n: the reference's square adjacency)."""
    n_cols = n if n_cols is None else int(n_cols)
    _require_cuda(row, "row")
    _require_cuda(col, "col")
    row = row.contiguous().to(torch.int64)
    col = col.contiguous().to(torch.int64)
    if row.shape != col.shape or row.dim() != 1:
        raise RuntimeError("row and col must be 1-D tensors of equal length")
    nnz = row.numel()
    dev = row.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        col_ind = torch.empty(nnz, dtype=torch.int32, device=dev)
        rows = torch.empty(nnz, dtype=torch.int32, device=dev)
        perm = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float32, device=dev)
        ws_bytes = int(L.dfgnn_format_workspace_bytes(max(n, n_cols), nnz))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        rc = L.dfgnn_coo_to_csr(n, n_cols, nnz, row.data_ptr() if nnz else None,
                                col.data_ptr() if nnz else None, row_ptr.data_ptr(),
                                col_ind.data_ptr() if nnz else None,
                                rows.data_ptr() if nnz else None,
                                perm.data_ptr() if nnz else None,
                                val.data_ptr() if nnz else None, ws.data_ptr(), ws_bytes,
                                torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "coo_to_csr")
    val._dfgnn_ones = True  # operators/_native.py: the kernels skip the weight loads for this tensor
    return row_ptr, col_ind, rows, perm, val


def csr_to_csc(row_ptr: torch.Tensor, col_ind: torch.Tensor, n_cols: int = None,
               rows: torch.Tensor = None):
    """-> col_ptr[n_cols+1], row_ind[E], val_idx[E] (all int32); val_idx = CSC pos -> CSR pos.
    n_cols defaults to the number of rows (square adjacency).  ``rows`` (the expanded row ids
    that coo_to_csr returns) is optional: with it row_ind is a gather instead of a search."""
    _require_cuda(row_ptr, "row_ptr")
    _require_cuda(col_ind, "col_ind")
    if row_ptr.dtype != torch.int32 or col_ind.dtype != torch.int32:
        raise RuntimeError("row_ptr and col_ind must be int32")
    row_ptr = row_ptr.contiguous()
    col_ind = col_ind.contiguous()
    n_rows = row_ptr.numel() - 1
    n = n_rows if n_cols is None else int(n_cols)
    nnz = col_ind.numel()
    dev = row_ptr.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        col_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        row_ind = torch.empty(nnz, dtype=torch.int32, device=dev)
        val_idx = torch.empty(nnz, dtype=torch.int32, device=dev)
        ws_bytes = int(L.dfgnn_format_workspace_bytes(max(n, n_rows), nnz))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        if rows is not None and (rows.dtype != torch.int32 or rows.numel() != nnz or not rows.is_contiguous()):
            raise RuntimeError("rows must be a contiguous int32 tensor of nnz entries")
        rc = L.dfgnn_csr_to_csc(n_rows, n, nnz, row_ptr.data_ptr(), col_ind.data_ptr() if nnz else None,
                                rows.data_ptr() if (rows is not None and nnz) else None,
                                col_ptr.data_ptr(), row_ind.data_ptr() if nnz else None,
                                val_idx.data_ptr() if nnz else None, ws.data_ptr(), ws_bytes,
                                torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "csr_to_csc")
    return col_ptr, row_ind, val_idx


class BlockPlan:
    """Node ranges of the graphs of a block-diagonal batch (``g.batch_num_nodes()``), validated
    against the CSR.  Attached to the ``row_ptr`` tensor the preprocessing returns
    (``row_ptr._dfgnn_blocks``): the operators then run the graph-resident kernels
    (csrc/block_gt.cuh) for the sizes they support."""

    # dense tensor-core kernels from this fill ratio of the diagonal blocks on (automatic mode)
    DENSE_MIN_FILL = 0.15
    # ... and, for the tcgen05 kernels, from this many edges per 128-row tile on: their cost is per TILE
    # (~54 us of one SM for forward + backward, whatever the tile holds) while the per-edge kernels cost
    # ~0.21 ns per edge of the whole GPU (measured on the PATTERN-shaped batch); batches of tiny graphs stay
    # on the per-edge kernels
    TC_MIN_EDGES_PER_TILE = 1800

    def __init__(self, blk_ptr: torch.Tensor, n_blocks: int, max_nodes: int, sum_sq_nodes: int = 0,
                 ascending: bool = True):
        self.blk_ptr = blk_ptr
        self.n_blocks = int(n_blocks)
        self.max_nodes = int(max_nodes)
        self.sum_sq_nodes = int(sum_sq_nodes)   # sum over graphs of nodes^2 = entries of the dense blocks
        self.ascending = bool(ascending)        # column ids strictly ascending inside every row
        self.adj_bits = None                    # [m, 8] int32 adjacency bitmap (dense tcgen05 kernels)
        self.n_ctas, self.sched_ptr, self.sched_idx = 0, None, None  # balanced graph lists of the persistent CTAs
        self.col_sched = (0, None, None)        # the same for the (graph, key tile) items of the column-side backward
        self.tile_ptr = None                    # [n_blocks + 1] first 128-row tile of every graph (dense work arrays)
        self.n_tiles = 0
        self._ok = {}

    def algorithm(self, m: int, nnz: int, h: int, f: int, unweighted: bool, training: bool = False) -> int:
        """0 general kernels, 1 shared-memory-staged sparse kernels, 2 dense mma.sync kernels
        (forward), 3 dense tcgen05 kernels (csrc/dense_tc.cu); see dfgnn_set_block_mode.  The dense
        kernels need unweighted scores and strictly ascending column ids per row (no duplicate
        edges).  Automatic mode picks the tcgen05 kernels for dense batches wherever they apply
        (f == 128, graphs of at most 256 nodes), else the mma.sync kernels for inference only."""
        L = _lib.lib()
        mode = L.dfgnn_set_block_mode(-1)
        key = ("algo", m, nnz, h, f, unweighted, training, mode)
        if key not in self._ok:
            algo = 0
            fill = nnz / self.sum_sq_nodes if self.sum_sq_nodes > 0 else 0.0
            dense_ok = unweighted and self.ascending and nnz > 0
            if (dense_ok and self.adj_bits is not None and mode in (0, 4)
                    and L.dfgnn_gt_dense_tc_supported(self.max_nodes, h, f)):
                if mode == 4 or self.prefers_dense_tc(nnz):
                    algo = 3
            if algo == 0 and dense_ok and L.dfgnn_gt_dense_supported(self.max_nodes, h, f):
                if mode == 3 or (fill >= self.DENSE_MIN_FILL and not training):
                    algo = 2
            if algo == 0 and L.dfgnn_gt_block_supported(self.max_nodes, m, nnz, h, f):
                algo = 1
            self._ok[key] = algo
        return self._ok[key]

    def prefers_dense_tc(self, nnz: int) -> bool:
        """Automatic mode: dense enough blocks AND enough edges per 128-row tile for the tcgen05 kernels."""
        fill = nnz / self.sum_sq_nodes if self.sum_sq_nodes > 0 else 0.0
        return fill >= self.DENSE_MIN_FILL and nnz >= self.TC_MIN_EDGES_PER_TILE * max(1, self.n_tiles)

    def supported(self, m: int, nnz: int, h: int, f: int) -> bool:
        key = (m, nnz, h, f, _lib.lib().dfgnn_set_block_mode(-1))
        if key not in self._ok:
            self._ok[key] = bool(_lib.lib().dfgnn_gt_block_supported(self.max_nodes, m, nnz, h, f))
        return self._ok[key]


def block_plan(batch_num_nodes: torch.Tensor, row_ptr: torch.Tensor, col_ind: torch.Tensor):
    """-> BlockPlan for a batch whose graphs have ``batch_num_nodes`` nodes each, or None when the
    batch is a single graph.  Raises if the CSR is not block diagonal over those ranges."""
    import ctypes
    _require_cuda(row_ptr, "row_ptr")
    bnn = batch_num_nodes.to(torch.int64).cpu()
    if bnn.numel() <= 1:
        return None
    m = row_ptr.numel() - 1
    if int(bnn.sum()) != m:
        raise RuntimeError(f"batch_num_nodes sums to {int(bnn.sum())}, the graph has {m} nodes")
    dev = row_ptr.device
    blk = torch.zeros(bnn.numel() + 1, dtype=torch.int32)
    blk[1:] = torch.cumsum(bnn, 0).to(torch.int32)
    blk = blk.to(dev)
    with torch.cuda.device(dev):
        flag = torch.empty(4, dtype=torch.int32, device=dev)
        mx, asc = ctypes.c_int32(0), ctypes.c_int32(0)
        rc = _lib.lib().dfgnn_block_plan_check(
            bnn.numel(), m, col_ind.numel(), blk.data_ptr(), row_ptr.data_ptr(),
            col_ind.data_ptr() if col_ind.numel() else None, flag.data_ptr(), ctypes.addressof(mx),
            ctypes.addressof(asc), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "block_plan")
    plan = BlockPlan(blk, bnn.numel(), mx.value, int((bnn * bnn).sum()), bool(asc.value))
    if plan.ascending and col_ind.numel() and mx.value <= 256:
        # adjacency bitmap for the dense tcgen05 kernels (32 bytes per row), built once per batch
        with torch.cuda.device(dev):
            bits = torch.empty((m, 8), dtype=torch.int32, device=dev)
            rc = _lib.lib().dfgnn_block_adj_bits(
                plan.n_blocks, plan.max_nodes, m, col_ind.numel(), blk.data_ptr(), row_ptr.data_ptr(),
                col_ind.data_ptr(), bits.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "block_adj_bits")
        plan.adj_bits = bits
        plan.n_ctas, plan.sched_ptr, plan.sched_idx = _balanced_schedule(bnn, dev)
        plan.col_sched = _balanced_schedule(bnn, dev, column_items=True)
        tp = torch.zeros(bnn.numel() + 1, dtype=torch.int32)
        tp[1:] = torch.cumsum((bnn + 127) // 128, 0).to(torch.int32)
        plan.tile_ptr = tp.to(dev)
        plan.n_tiles = int(tp[-1])
    return plan


def balanced_lists(bnn, n_ctas: int, column_items: bool = False):
    """Host part of ``_balanced_schedule``: longest-processing-time-first lists from the graph sizes
    (``dfgnn_tc_balanced_lists``, plain C++ on the host: the plan is built per batch, a Python heap loop
    over 2400 items cost more than the conv step).  -> (ctas, ptr [ctas + 1] int32, ids int32)."""
    import numpy as np
    n = np.ascontiguousarray(np.asarray(bnn, dtype=np.int32))
    ptr = np.zeros(int(n_ctas) + 1, dtype=np.int32)
    idx = np.zeros(2 * len(n), dtype=np.int32)
    g = _lib.lib().dfgnn_tc_balanced_lists(len(n), n.ctypes.data, int(n_ctas), int(bool(column_items)),
                                           ptr.ctypes.data, idx.ctypes.data)
    if g < 1:
        raise RuntimeError("dfgnn_tc_balanced_lists: invalid arguments")
    return g, ptr[: g + 1].copy(), idx[: int(ptr[g])].copy()


def _balanced_schedule(bnn: torch.Tensor, dev, column_items: bool = False):
    """Work lists for the persistent CTAs of the dense tcgen05 kernels (one CTA per SM): longest
    processing time first over a cost model of the MMA work.  Forward / row side: one entry per
    graph (a graph of more than 128 nodes is two row tiles over up to 256 keys).  Column side
    (``column_items``): one entry per (graph, 128-key tile), id = 2 * graph + tile, cost = the 16-row
    slices of the graph.  Dealing 1024 PATTERN-shaped graphs round robin leaves the busiest SM with
    1.65x the mean work; this schedule 1.05x.  -> (number of CTAs, ptr [ctas + 1], ids)."""
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    g, ptr, idx = balanced_lists(bnn, sms, column_items)
    return g, torch.from_numpy(ptr).to(dev), torch.from_numpy(idx).to(dev)


def attach_block_plan(g, row_ptr: torch.Tensor, col_ind: torch.Tensor):
    """Preprocessing hook: a batched graph (``g.batch_size > 1``) gets its block plan attached to
    ``row_ptr``.  Graphs without batch information are left alone."""
    try:
        bs = int(getattr(g, "batch_size", 1))
        if bs <= 1 or getattr(g, "num_cols", g.num_nodes()) != g.num_nodes():
            return
        bnn = g.batch_num_nodes()
    except Exception:
        return
    plan = block_plan(bnn, row_ptr, col_ind)
    if plan is not None:
        row_ptr._dfgnn_blocks = plan
        _register_plan(row_ptr, plan)


# Tensors that come back from autograd's save_for_backward are new Python objects over the same
# memory: the plan is also findable by address for as long as the tensor it was attached to lives
# (its memory cannot be reused before that).
_PLANS = {}


def _register_plan(row_ptr: torch.Tensor, plan) -> None:
    import weakref
    dead = [k for k, (ref, _) in _PLANS.items() if ref() is None]
    for k in dead:
        del _PLANS[k]
    _PLANS[(row_ptr.data_ptr(), row_ptr.numel())] = (weakref.ref(row_ptr), plan)


def find_block_plan(row_ptr: torch.Tensor):
    plan = getattr(row_ptr, "_dfgnn_blocks", None)
    if plan is not None:
        return plan
    hit = _PLANS.get((row_ptr.data_ptr(), row_ptr.numel()))
    if hit is None:
        return None
    owner = hit[0]()
    if owner is None or owner.data_ptr() != row_ptr.data_ptr() or owner.device != row_ptr.device:
        return None
    return hit[1]
