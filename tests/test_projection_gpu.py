"""The tcgen05 projection (csrc/proj_tc.cu) against torch fp64: q/k/v of SparseMHA.prep_qkv
(DFGNN/layers/GT/gtconv_layer.py:19-27) and the GAT prologue (gatconv_layer_fused.py:121-123,
fused_gatconv_hyper_v2.cu:212-250).  Tolerance 1e-4 relative / 1e-5 absolute: the kernel is
fp32-grade (3xTF32), not single-pass TF32 (which misses this bar by ~10x; checked below)."""
import pytest
import torch

from dfgnn_b200.operators.projection import (FusedGATProjFunction, FusedQKVFunction, PackedWeights,
                                             proj_forward, supported)

from .helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k,d,heads", [(1000, 128, 128, 1), (4099, 64, 64, 2), (129, 32, 64, 1), (128, 128, 192, 3),
                                         (1, 64, 64, 1), (70000, 128, 128, 1)])
def test_fused_qkv_matches_fp64(cuda, n, k, d, heads):
    assert supported(k, d, 3)
    torch.manual_seed(n)
    x = torch.randn(n, k, device=cuda)
    lin = [torch.nn.Linear(k, d).to(cuda) for _ in range(3)]
    scaling = (d // heads) ** -0.5
    cache = PackedWeights()
    q, kk, v = FusedQKVFunction.apply(x, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight,
                                      lin[2].bias, scaling, heads, cache)
    assert q.shape == kk.shape == v.shape == (n, heads, d // heads) and q.is_contiguous()
    xd = x.double()
    ref = [(xd @ l.weight.double().t() + l.bias.double()) for l in lin]
    assert_close("q", q.reshape(n, d), ref[0] * scaling)
    assert_close("k", kk.reshape(n, d), ref[1])
    assert_close("v", v.reshape(n, d), ref[2])
    # a second call reuses the packed images; a weight update rebuilds them
    img0 = cache.img
    FusedQKVFunction.apply(x, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight, lin[2].bias,
                           scaling, heads, cache)
    assert cache.img is img0
    with torch.no_grad():
        lin[1].weight.mul_(2.0)
    _, k2, _ = FusedQKVFunction.apply(x, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight,
                                      lin[2].bias, scaling, heads, cache)
    assert cache.img is not img0
    assert_close("k after the weight update", k2.reshape(n, d), xd @ lin[1].weight.double().t() + lin[1].bias.double())


def test_single_pass_tf32_would_not_meet_the_bar(cuda):
    """Why the split: torch's own TF32 GEMM on the same inputs misses 1e-4 / 1e-5."""
    torch.manual_seed(0)
    x = torch.randn(4096, 128, device=cuda)
    w = torch.randn(128, 128, device=cuda) * 128 ** -0.5
    ref = x.double() @ w.double().t()
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        y = x @ w.t()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False
    bad = ((y.double() - ref).abs() > 1e-5 + 1e-4 * ref.abs()).float().mean()
    assert float(bad) > 0.05


@pytest.mark.parametrize("n,k,heads,d", [(5000, 128, 1, 64), (3001, 64, 1, 128), (777, 64, 4, 16), (640, 32, 8, 8),
                                         (900, 128, 2, 32)])
def test_fused_gat_projection_and_logits(cuda, n, k, heads, d):
    torch.manual_seed(3)
    x = torch.randn(n, k, device=cuda)
    W = torch.nn.Linear(k, heads * d, bias=False).to(cuda)
    a_l = torch.randn(1, heads, d, device=cuda)
    a_r = torch.randn(1, heads, d, device=cuda)
    feat, ar, ac = FusedGATProjFunction.apply(x, W.weight, None, a_l, a_r, heads, PackedWeights())
    ref = (x.double() @ W.weight.double().t()).view(n, heads, d)
    assert_close("feat", feat, ref)
    assert_close("attn_row", ar, (a_l.double() * ref).sum(-1))
    assert_close("attn_col", ac, (a_r.double() * ref).sum(-1))


def test_projection_functions_train_like_linear(cuda):
    torch.manual_seed(5)
    n, k, d = 2000, 64, 64
    x = torch.randn(n, k, device=cuda, requires_grad=True)
    lin = [torch.nn.Linear(k, d).to(cuda) for _ in range(3)]
    q, kk, v = FusedQKVFunction.apply(x, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight,
                                      lin[2].bias, 0.25, 1, PackedWeights())
    wq, wk, wv = (torch.randn_like(t) for t in (q, kk, v))
    ((q * wq).sum() + (kk * wk).sum() + (v * wv).sum()).backward()
    got = [x.grad.clone()] + [l.weight.grad.clone() for l in lin] + [l.bias.grad.clone() for l in lin]
    x.grad = None
    for l in lin:
        l.zero_grad()
    q2 = (lin[0](x) * 0.25).view(n, 1, d)
    k2, v2 = lin[1](x).view(n, 1, d), lin[2](x).view(n, 1, d)
    ((q2 * wq).sum() + (k2 * wk).sum() + (v2 * wv).sum()).backward()
    want = [x.grad] + [l.weight.grad for l in lin] + [l.bias.grad for l in lin]
    for i, (a, b) in enumerate(zip(got, want)):
        assert_close(f"grad {i}", a, b, rtol=1e-3, atol=1e-3)
    # GAT
    W = torch.nn.Linear(k, 64, bias=False).to(cuda)
    a_l = torch.randn(1, 1, 64, device=cuda, requires_grad=True)
    a_r = torch.randn(1, 1, 64, device=cuda, requires_grad=True)
    xg = x.detach().requires_grad_()
    feat, ar, ac = FusedGATProjFunction.apply(xg, W.weight, None, a_l, a_r, 1, PackedWeights())
    wf, w1, w2 = torch.randn_like(feat), torch.randn_like(ar), torch.randn_like(ac)
    ((feat * wf).sum() + (ar * w1).sum() + (ac * w2).sum()).backward()
    got = [xg.grad.clone(), W.weight.grad.clone(), a_l.grad.clone(), a_r.grad.clone()]
    xg.grad = None
    W.zero_grad()
    a_l.grad = a_r.grad = None
    f2 = W(xg).view(n, 1, 64)
    ((f2 * wf).sum() + ((a_l * f2).sum(-1) * w1).sum() + ((a_r * f2).sum(-1) * w2).sum()).backward()
    for i, (a, b) in enumerate(zip(got, [xg.grad, W.weight.grad, a_l.grad, a_r.grad])):
        assert_close(f"gat grad {i}", a, b, rtol=1e-3, atol=1e-3)
