"""Format preprocessing and layer factories with the reference's names and return
tuples (``DFGNN/layers/util.py``).  ``g`` is anything with ``edges()`` and
``num_nodes()`` -- a ``dgl.DGLGraph`` or ``dfgnn_b200.graphs.Graph`` -- living on
the GPU; all index work runs in the CUDA library (``dfgnn_b200/formats.py``)."""
from __future__ import annotations

import torch

from ..formats import SparseMatrix, attach_block_plan, coo_to_csr, csr_to_csc

WARP_SIZE = 32


def g_to_SPmatrix(g):
    """layers/util.py:52-57: COO of the graph + the hard-coded max_neigh = 128."""
    row, col = g.edges()
    N = g.num_nodes()
    return SparseMatrix(row, col, (N, getattr(g, "num_cols", N))), 128


def preprocess_dglsp(g, **args):
    """DFGNN/utils/util.py:239-243: operand of the non-fused branch."""
    return g_to_SPmatrix(g)[0]


def _smem(max_neigh: int, mult: int) -> int:
    return (max_neigh * mult + WARP_SIZE - 1) // WARP_SIZE * WARP_SIZE


def preprocess_CSR(g, **args):
    """layers/util.py:66-79 -> (row_ptr, col_ind, val, smem_consume=128)."""
    A, max_neigh = g_to_SPmatrix(g)
    row_ptr, col_ind, _, _, val = coo_to_csr(A.row, A.col, A.shape[0], A.shape[1])
    attach_block_plan(g, row_ptr, col_ind)
    return row_ptr, col_ind, val, _smem(max_neigh, 1)


def preprocess_Hyper(g, **args):
    """layers/util.py:82-100 -> (row_ptr, col_ind, rows, val, smem_consume=1024)."""
    A, max_neigh = g_to_SPmatrix(g)
    row_ptr, col_ind, rows, _, val = coo_to_csr(A.row, A.col, A.shape[0], A.shape[1])
    attach_block_plan(g, row_ptr, col_ind)
    return row_ptr, col_ind, rows, val, _smem(max_neigh, 8)


def preprocess_softmax(g, **args):
    """layers/util.py:145-162 -> (row_ptr, col_ind, rows, val, smem_consume=128)."""
    A, max_neigh = g_to_SPmatrix(g)
    row_ptr, col_ind, rows, _, val = coo_to_csr(A.row, A.col, A.shape[0], A.shape[1])
    attach_block_plan(g, row_ptr, col_ind)
    return row_ptr, col_ind, rows, val, _smem(max_neigh, 1)


def preprocess_Hyper_fw_bw(g, fused=True):
    """layers/util.py:116-142 -> (A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx,
    smem_consume)."""
    A, max_neigh = g_to_SPmatrix(g)
    if not fused:
        return A, None, None, None, None, None, None, None, None
    row_ptr, col_ind, rows, _, val = coo_to_csr(A.row, A.col, A.shape[0], A.shape[1])
    col_ptr, row_ind, val_idx = csr_to_csc(row_ptr, col_ind, A.shape[1], rows)
    attach_block_plan(g, row_ptr, col_ind)
    return A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, _smem(max_neigh, 8)


def preprocess_gat_fw_bw(g):
    """CSR + CSC + permute for FusedGATFunction (what
    DFGNN/script/train/train_gatconv.py:119-136 builds with scipy).
    -> (row_ptr, col_ind, col_ptr, row_ind, permute)."""
    A, _ = g_to_SPmatrix(g)
    row_ptr, col_ind, rows, _, _ = coo_to_csr(A.row, A.col, A.shape[0], A.shape[1])
    col_ptr, row_ind, permute = csr_to_csc(row_ptr, col_ind, A.shape[1], rows)
    return row_ptr, col_ind, col_ptr, row_ind, permute


def load_layer_GT(args):
    """layers/util.py:362-393 (formats whose kernels are on the hot path)."""
    from .GT import (SparseMHA_CSR, SparseMHA_CSR_GM, SparseMHA_forward_timing, SparseMHA_hyper,
                     SparseMHA_softmax, SparseMHA_softmax_gm, SparseMHA_tiling)
    table = {"csr": SparseMHA_CSR, "csr_gm": SparseMHA_CSR_GM, "tiling": SparseMHA_tiling,
             "hyper": SparseMHA_hyper, "nofuse": SparseMHA_hyper, "softmax": SparseMHA_softmax,
             "softmax_gm": SparseMHA_softmax_gm, "forward": SparseMHA_forward_timing}
    if args.format not in table:
        raise ValueError(f"Unsupported format {args.format} in GTconv")
    return table[args.format](args.dim, args.dim, args.heads)


def load_layer_GAT(args):
    """layers/util.py:396-421."""
    from .GAT import (GATConv_dgNN, GATConv_hyper, GATConv_hyper_recompute, GATConv_hyper_v2,
                      GATConv_softmax, GATConv_softmax_gm, GATConv_tiling)
    table = {"csr": GATConv_dgNN, "tiling": GATConv_tiling, "hyper": GATConv_hyper,
             "nofuse": GATConv_hyper, "hyper_v2": GATConv_hyper_v2,
             "hyper_recompute": GATConv_hyper_recompute, "softmax": GATConv_softmax,
             "softmax_gm": GATConv_softmax_gm}
    if args.format not in table:
        raise ValueError(f"Unsupported format {args.format} in GATconv")
    return table[args.format](args.dim, args.dim, args.heads)


def load_layer_AGNN(args):
    """layers/util.py:424-443."""
    from .AGNN import (AGNNConv_csr, AGNNConv_csr_gm, AGNNConv_hyper, AGNNConv_softmax,
                       AGNNConv_softmax_gm, AGNNConv_tiling)
    table = {"hyper": AGNNConv_hyper, "nofuse": AGNNConv_hyper, "csr": AGNNConv_csr,
             "softmax": AGNNConv_softmax, "csr_gm": AGNNConv_csr_gm, "tiling": AGNNConv_tiling,
             "softmax_gm": AGNNConv_softmax_gm}
    if args.format not in table:
        raise ValueError(f"Unsupported format {args.format} in AGNNconv")
    return table[args.format](args.dim, args.dim, args.heads)


def load_graphconv_layer(args):
    """layers/util.py:446-455."""
    if args.conv == "gat":
        return load_layer_GAT(args)
    if args.conv == "gt":
        return load_layer_GT(args)
    if args.conv == "agnn":
        return load_layer_AGNN(args)
    raise ValueError(f"unknown graph conv {args.conv}")


def load_prepfunc(args):
    """layers/util.py:458-491."""
    if args.format in ["csr", "csr_gm", "tiling"]:
        return preprocess_CSR
    if args.format in ["hyper", "nofuse", "hyper_ablation", "hyper_recompute", "hyper_v2"]:
        return preprocess_Hyper
    if args.format in ["softmax", "softmax_gm"]:
        return preprocess_softmax
    if args.format == "forward":
        return preprocess_Hyper_fw_bw
    raise ValueError(f"Unsupported format {args.format}")
