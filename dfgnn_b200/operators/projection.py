"""The dense projection in front of the conv on the tcgen05 tensor cores (csrc/proj_tc.cu).

Mirror of ``SparseMHA.prep_qkv`` + the layout transposes (``DFGNN/layers/GT/gtconv_layer.py:19-27``,
``gtconv_layer_fused.py:20-22``) and of the GAT prologue ``feat = W(x)``, ``attn_row / attn_col``
(``layers/GAT/gatconv_layer_fused.py:121-123``): one kernel instead of three fp32 cuBLAS GEMMs, a
scale, transposes and two reductions, at fp32-grade accuracy (3xTF32).  The backward of the linear
maps is plain torch (two GEMMs per map), so the Functions below train like ``nn.Linear``."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .. import _lib
from ._native import _chk, _ptr, _stream

_SUPPORTED_K = (32, 64, 128)


def supported(k: int, part_width: int, parts: int) -> bool:
    return k in _SUPPORTED_K and part_width % 64 == 0 and 1 <= parts <= 4


class PackedWeights:
    """hi / lo operand images of a stack of weight matrices [sum(out_i), k] (rows of the parts
    concatenated), rebuilt when any of the weights changes (``Tensor._version``)."""

    def __init__(self):
        self.key = None
        self.img = None
        self.bias = None

    def get(self, weights: Sequence[torch.Tensor], biases: Sequence[Optional[torch.Tensor]],
            row_perm: Optional[torch.Tensor] = None):
        """``row_perm``: output-column permutation applied to every part (row i of the packed part is
        row row_perm[i] of the weight) -- the reference's inference modules split the projection
        output as [N, head_dim, heads] and transpose (gtconv_layer.py:19-27, gtconv_layer_fused.py:
        20-22); permuting the weight rows yields [N, heads, head_dim] directly."""
        key = tuple((w.data_ptr(), w._version, tuple(w.shape)) for w in weights) + \
            tuple((b.data_ptr(), b._version) if b is not None else None for b in biases) + \
            (None if row_perm is None else int(row_perm.numel()),)
        if key != self.key:
            if row_perm is not None:
                weights = [w[row_perm] for w in weights]
                biases = [b[row_perm] if b is not None else None for b in biases]
            W = torch.cat([w.detach() for w in weights], 0).contiguous().float()
            n_out, k = W.shape
            L = _lib.lib()
            with torch.cuda.device(W.device):
                img = torch.empty(int(L.dfgnn_proj_weight_image_floats(n_out, k)), dtype=torch.float32, device=W.device)
                rc = L.dfgnn_proj_pack_weights(n_out, k, W.data_ptr(), img.data_ptr(), _stream(W))
            _lib.check(rc, "proj_pack_weights")
            self.img = img
            self.bias = None
            if any(b is not None for b in biases):
                self.bias = torch.cat([b.detach() if b is not None else W.new_zeros(w.shape[0])
                                       for w, b in zip(weights, biases)]).contiguous()
            self.key = key
        return self.img, self.bias


def proj_forward(x: torch.Tensor, img: torch.Tensor, bias, scale, n_out: int, part_width: int,
                 head_dim: int = 0, a_l=None, a_r=None):
    """-> (list of n_out / part_width tensors [n, part_width], attn_row, attn_col)."""
    fn = "proj_forward"
    _chk("x", x, torch.float32)
    n, k = x.shape
    parts = n_out // part_width
    dev = x.device
    with torch.cuda.device(dev):
        outs = [torch.empty((n, part_width), dtype=torch.float32, device=dev) for _ in range(parts)]
        ar = ac = None
        if head_dim > 0:
            ar = torch.empty((n, n_out // head_dim), dtype=torch.float32, device=dev)
            ac = torch.empty_like(ar)
        op = [_ptr(o) if o.numel() else o.data_ptr() for o in outs] + [None] * (4 - parts)
        rc = _lib.lib().dfgnn_proj_forward(
            n, k, n_out, part_width, _ptr(x), _ptr(img), _ptr(bias), _ptr(scale), op[0], op[1], op[2], op[3],
            int(head_dim), _ptr(a_l), _ptr(a_r), _ptr(ar), _ptr(ac), _stream(x))
    _lib.check(rc, fn)
    return outs, ar, ac


class FusedQKVFunction(torch.autograd.Function):
    """q, k, v = ((x Wq^T + bq) * scaling, x Wk^T + bk, x Wv^T + bv), each [N, heads, d]."""

    @staticmethod
    def forward(ctx, x, wq, bq, wk, bk, wv, bv, scaling: float, heads: int, cache: PackedWeights,
                row_perm=None):
        img, bias = cache.get((wq, wk, wv), (bq, bk, bv), row_perm)
        ctx.row_perm = row_perm
        d_out = wq.shape[0]
        scale = torch.ones(3 * d_out, dtype=torch.float32, device=x.device)
        scale[:d_out] = scaling
        (q, k, v), _, _ = proj_forward(x.contiguous(), img, bias, scale, 3 * d_out, d_out)
        ctx.save_for_backward(x, wq, wk, wv)
        ctx.scaling, ctx.has_bias = scaling, (bq is not None, bk is not None, bv is not None)
        shape = (x.shape[0], heads, d_out // heads)
        return q.view(shape), k.view(shape), v.view(shape)

    @staticmethod
    def backward(ctx, gq, gk, gv):
        x, wq, wk, wv = ctx.saved_tensors
        n = x.shape[0]
        gq = gq.reshape(n, -1) * ctx.scaling
        gk, gv = gk.reshape(n, -1), gv.reshape(n, -1)
        if ctx.row_perm is not None:  # packed column i is weight row row_perm[i]
            inv = torch.empty_like(ctx.row_perm)
            inv[ctx.row_perm] = torch.arange(ctx.row_perm.numel(), device=inv.device)
            gq, gk, gv = gq[:, inv], gk[:, inv], gv[:, inv]
        gx = gq @ wq + gk @ wk + gv @ wv if ctx.needs_input_grad[0] else None
        gb = [g.sum(0) if hb else None for g, hb in zip((gq, gk, gv), ctx.has_bias)]
        return gx, gq.t() @ x, gb[0], gk.t() @ x, gb[1], gv.t() @ x, gb[2], None, None, None, None


class FusedGATProjFunction(torch.autograd.Function):
    """feat = x W^T (+ b) as [N, heads, d]; attn_row = <a_l, feat>, attn_col = <a_r, feat> [N, heads]."""

    @staticmethod
    def forward(ctx, x, w, b, a_l, a_r, heads: int, cache: PackedWeights):
        img, bias = cache.get((w,), (b,))
        n_out = w.shape[0]
        d = n_out // heads
        (feat,), ar, ac = proj_forward(x.contiguous(), img, bias, None, n_out, n_out, d,
                                       a_l.detach().reshape(-1).contiguous(), a_r.detach().reshape(-1).contiguous())
        ctx.save_for_backward(x, w, a_l, a_r, feat)
        ctx.heads, ctx.has_bias = heads, b is not None
        return feat.view(x.shape[0], heads, d), ar, ac

    @staticmethod
    def backward(ctx, gfeat, gar, gac):
        x, w, a_l, a_r, feat = ctx.saved_tensors
        n, heads = x.shape[0], ctx.heads
        d = w.shape[0] // heads
        al, ar_ = a_l.reshape(1, heads, d), a_r.reshape(1, heads, d)
        f3 = feat.view(n, heads, d)
        g = gfeat.reshape(n, heads, d) + gar.unsqueeze(-1) * al + gac.unsqueeze(-1) * ar_
        g2 = g.reshape(n, -1)
        gx = g2 @ w if ctx.needs_input_grad[0] else None
        gal = (gar.unsqueeze(-1) * f3).sum(0).reshape(a_l.shape)
        gar_ = (gac.unsqueeze(-1) * f3).sum(0).reshape(a_r.shape)
        return gx, g2.t() @ x, (g2.sum(0) if ctx.has_bias else None), gal, gar_, None, None


def fused_qkv(x, q_proj, k_proj, v_proj, scaling: float, heads: int, cache: PackedWeights,
              interleaved_heads: bool = False):
    """The three ``nn.Linear`` modules of SparseMHA in one tensor-core kernel -> q, k, v [N, heads, d].
    ``interleaved_heads``: the projection output is split as [N, d, heads] (prep_qkv) rather than
    [N, heads, d] (the training module); same values, transposed for the conv."""
    perm = None
    d_out = q_proj.weight.shape[0]
    if interleaved_heads and heads > 1:
        hd = d_out // heads
        o = torch.arange(d_out, device=x.device)
        perm = (o % hd) * heads + o // hd     # packed column h * hd + d  <-  weight row d * heads + h
    return FusedQKVFunction.apply(x, q_proj.weight, q_proj.bias, k_proj.weight, k_proj.bias,
                                  v_proj.weight, v_proj.bias, scaling, heads, cache, perm)
