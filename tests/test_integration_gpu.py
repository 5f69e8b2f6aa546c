"""The reference's UNMODIFIED operator files running on the B200 kernels.

``DFGNN/operators/fused_gtconv.py`` / ``fused_gatconv.py`` do ``import fused_gtconv as fused_gt``
/ ``import fused_gatconv as fused_gat`` (operators/fused_gtconv.py:1, fused_gatconv.py:1).  With
``integration/`` first on ``sys.path`` those imports resolve to the shim modules, i.e. to
libdfgnn_b200.so through ctypes, and the reference's own autograd Functions
(operators/fused_gtconv.py:79-158, fused_gatconv.py:95-176) drive our forward and backward.

The operator files are byte-for-byte copies staged by oracle/build_ref.py under the git-ignored
oracle/_ref/operators/ (the GPU box has no /root/reference); skipped when they are absent."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from dfgnn_b200 import _lib, graphs
from dfgnn_b200.layers import preprocess_gat_fw_bw, preprocess_Hyper_fw_bw
from oracle import cpu_oracle as O

from .helpers import assert_close

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = os.path.join(ROOT, "oracle", "_ref", "operators")
needs_ops = pytest.mark.skipif(
    not all(os.path.exists(os.path.join(OPS, f)) for f in ("fused_gtconv.py", "fused_gatconv.py")),
    reason="oracle/_ref/operators not staged (run oracle/build_ref.py where /root/reference exists)")


def _load_reference_operator(fname):
    """Import oracle/_ref/operators/<fname> with integration/ shadowing the pybind modules."""
    shim_dir = os.path.join(ROOT, "integration")
    saved = {m: sys.modules.pop(m, None) for m in ("fused_gtconv", "fused_gatconv")}
    sys.path.insert(0, shim_dir)
    try:
        spec = importlib.util.spec_from_file_location("refop_" + fname[:-3], os.path.join(OPS, fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        backend = sys.modules[fname[:-3]]
        assert os.path.dirname(os.path.abspath(backend.__file__)) == shim_dir, "the shim was not picked up"
    finally:
        sys.path.remove(shim_dir)
        for m, v in saved.items():
            sys.modules.pop(m, None)
            if v is not None:
                sys.modules[m] = v
    return mod


@needs_ops
def test_unmodified_reference_gt_operators_run_on_our_kernels(cuda):
    ref_ops = _load_reference_operator("fused_gtconv.py")
    g = graphs.pattern_like(batch=4)
    n = g.num_nodes()
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(cuda))
    X = graphs.conv_inputs(n, 128, 41)
    Q, K, V = (t.to(cuda).requires_grad_() for t in (X.Q, X.K, X.V))
    n0 = _lib.launch_count()
    out = ref_ops.GTConvFuse_hyper(rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    assert type(out.grad_fn).__name__ == "FusedGTFunction_hyperBackward"   # the reference's Function
    out.backward(X.dO.to(cuda))
    assert _lib.launch_count() - n0 >= 3, "libdfgnn_b200.so did not launch the kernels"
    rp, ci = row_ptr.cpu().numpy(), col_ind.cpu().numpy()
    cp, ri, vi = col_ptr.cpu().numpy(), row_ind.cpu().numpy(), val_idx.cpu().numpy()
    o64, a64 = O.gt_forward(rp, ci, None, X.Q, X.K, X.V, dtype=np.float64)
    dQ, dK, dV, _ = O.gt_backward(rp, ci, cp, ri, vi, X.Q, X.K, X.V, a64, X.dO, dtype=np.float64)
    assert_close("out", out, o64)
    assert_close("dQ", Q.grad, dQ)
    assert_close("dK", K.grad, dK)
    assert_close("dV", V.grad, dV)
    # every inference wrapper of the reference file
    Qd, Kd, Vd = Q.detach(), K.detach(), V.detach()
    for name, args in (("GTConvFuse_inference_hyper", (row_ptr, col_ind, rows, val, smem)),
                       ("GTConvFuse_inference_softmax", (row_ptr, col_ind, rows, val, 128)),
                       ("GTConvFuse_inference_softmax_gm", (row_ptr, col_ind, rows, val)),
                       ("GTConvFuse_inference_csr", (row_ptr, col_ind, val, 128)),
                       ("GTConvFuse_inference_csr_gm", (row_ptr, col_ind, val)),
                       ("GTConvFuse_inference_tiling", (row_ptr, col_ind, val, 128)),
                       ("GTConvFuse_inference_hyper_ablation", (row_ptr, col_ind, rows, val, smem))):
        assert_close(name, getattr(ref_ops, name)(*args, Qd, Kd, Vd), o64)


@needs_ops
def test_unmodified_reference_gat_operators_run_on_our_kernels(cuda):
    ref_ops = _load_reference_operator("fused_gatconv.py")
    g = graphs.arxiv_like(0.03)
    n = g.num_nodes()
    gd = g.to(cuda)
    row_ptr, col_ind, col_ptr, row_ind, permute = preprocess_gat_fw_bw(gd)
    X = graphs.conv_inputs(n, 64, 42)
    ar, ac, F = (t.to(cuda).requires_grad_() for t in (X.attn_row, X.attn_col, X.V))
    n0 = _lib.launch_count()
    out = ref_ops.GATConvFuse(ar, ac, row_ptr, col_ind, col_ptr, row_ind, permute, 0.2, F, 0.0)
    assert type(out.grad_fn).__name__ == "FusedGATFunctionBackward"
    out.backward(X.dO.to(cuda))
    assert _lib.launch_count() - n0 >= 3
    rp, ci = row_ptr.cpu().numpy(), col_ind.cpu().numpy()
    cp, ri, vi = col_ptr.cpu().numpy(), row_ind.cpu().numpy(), permute.cpu().numpy()
    o64, emax, esum = O.gat_forward(X.attn_row, X.attn_col, rp, ci, 0.2, X.V, dtype=np.float64)
    gf, gr, gc = O.gat_backward(0.2, 0.0, rp, ci, cp, ri, vi, emax, esum, None, X.V, X.attn_row,
                                X.attn_col, X.dO, dtype=np.float64)
    assert_close("out", out, o64)
    assert_close("d feat", F.grad, gf)
    assert_close("d attn_row", ar.grad, gr)
    assert_close("d attn_col", ac.grad, gc)
    rows = torch.repeat_interleave(torch.arange(n, device=cuda, dtype=torch.int32),
                                   (row_ptr[1:] - row_ptr[:-1]).long())
    ard, acd, Fd = ar.detach(), ac.detach(), F.detach()
    for name, args in (("GATConvFuse_inference", (ard, acd, row_ptr, col_ind, 0.2, Fd)),
                       ("GATConvFuse_inference_tiling", (ard, acd, row_ptr, col_ind, 0.2, Fd)),
                       ("GATConvFuse_inference_hyper", (1024, ard, acd, row_ptr, col_ind, rows, 0.2, Fd)),
                       ("GATConvFuse_inference_hyper_recompute", (ard, acd, row_ptr, col_ind, 0.2, Fd)),
                       ("GATConvFuse_inference_softmax", (128, ard, acd, row_ptr, col_ind, rows, 0.2, Fd)),
                       ("GATConvFuse_inference_softmax_gm", (ard, acd, row_ptr, col_ind, rows, 0.2, Fd))):
        assert_close(name, getattr(ref_ops, name)(*args), o64)
