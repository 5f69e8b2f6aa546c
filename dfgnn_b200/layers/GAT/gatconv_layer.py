"""GAT conv modules: names, parameters and ``forward(params, feat, fuse)`` contract of
``DFGNN/layers/GAT/*.py``; fused branches run the B200 kernels."""
import torch
from torch import nn

from ...operators.fused_gatconv import (GATConvFuse, GATConvFuse_inference,
                                        GATConvFuse_inference_hyper,
                                        GATConvFuse_inference_hyper_recompute,
                                        GATConvFuse_inference_hyper_v2,
                                        GATConvFuse_inference_softmax,
                                        GATConvFuse_inference_softmax_gm,
                                        GATConvFuse_inference_tiling)
from ...operators import projection
from ...utils import benchmark
from .._dglsp import bspmm, edge_softmax


class GATConvDGL(nn.Module):
    """layers/GAT/gatconv_layer.py:6-38."""

    def __init__(self, in_size, out_size, num_heads, dropout=0, negative_slope=0.2):
        super().__init__()
        self.in_size = in_size
        self.out_size = out_size
        self.num_heads = num_heads
        self.negative_slope = negative_slope
        self.dropout = nn.Dropout(dropout)
        self.W = nn.Linear(in_size, out_size * num_heads)
        self.a_l = nn.Parameter(torch.zeros(1, out_size, num_heads))
        self.a_r = nn.Parameter(torch.zeros(1, out_size, num_heads))
        self.activation = nn.LeakyReLU(negative_slope=negative_slope)
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.W.weight, gain=gain)
        nn.init.xavier_normal_(self.a_l, gain=gain)
        nn.init.xavier_normal_(self.a_r, gain=gain)

    def forward_dglsp(self, A_hat, Z):
        """gatconv_layer.py:30-38; Z is [N, out, heads]."""
        e_l = (Z * self.a_l).sum(dim=1)
        e_r = (Z * self.a_r).sum(dim=1)
        e = e_l[A_hat.row.long()] + e_r[A_hat.col.long()]
        a = self.activation(e)
        return bspmm(A_hat, edge_softmax(A_hat, a), Z)

    def _attn(self, a_l, a_r, h):
        """gatconv_layer_fused.py:121-123."""
        return (a_l * h).sum(dim=-1), (a_r * h).sum(dim=-1)

    def _forward(self, fused_conv, conv_args, params, feat, fuse):
        N = len(feat)
        if fuse:
            feat = self.W(feat).view(-1, self.num_heads, self.out_size)
            feat = feat.detach().contiguous()
            out, elapsed_time = benchmark(fused_conv, *conv_args, self.a_l.transpose(1, 2),
                                          self.a_r.transpose(1, 2), feat)
        else:
            feat = self.W(feat).view(-1, self.out_size, self.num_heads).detach().contiguous()
            out, elapsed_time = benchmark(self.forward_dglsp, params, feat)
        return out.reshape(N, -1), elapsed_time * 1000


class GATConv_dgNN(GATConvDGL):
    """gatconv_layer_fused.py:13-44 ("csr" format); params = preprocess_CSR(g)."""

    def conv(self, row_ptr, col_ind, a_l, a_r, h):
        attn_row, attn_col = self._attn(a_l, a_r, h)
        return GATConvFuse_inference(attn_row, attn_col, row_ptr, col_ind, self.negative_slope, h)

    def forward(self, params, feat, fuse=False):
        args = params[:2] if fuse else ()
        return self._forward(self.conv, args, params, feat, fuse)


class GATConv_tiling(GATConvDGL):
    """gatconv_layer_tiling.py:7-38; params = preprocess_CSR(g)."""

    def conv(self, row_ptr, col_ind, a_l, a_r, h):
        attn_row, attn_col = self._attn(a_l, a_r, h)
        return GATConvFuse_inference_tiling(attn_row, attn_col, row_ptr, col_ind,
                                            self.negative_slope, h)

    def forward(self, params, feat, fuse=False):
        args = params[:2] if fuse else ()
        return self._forward(self.conv, args, params, feat, fuse)


class GATConv_hyper(GATConvDGL):
    """gatconv_layer_fused.py:47-86; params = preprocess_Hyper(g)."""

    def conv(self, indptr, indices, rows, smem_consume, a_l, a_r, h):
        attn_row, attn_col = self._attn(a_l, a_r, h)
        return GATConvFuse_inference_hyper(smem_consume, attn_row, attn_col, indptr, indices, rows,
                                           self.negative_slope, h)

    def forward(self, params, feat, fuse=False):
        args = (params[0], params[1], params[2], params[4]) if fuse else ()
        return self._forward(self.conv, args, params, feat, fuse)


class GATConv_hyper_recompute(GATConvDGL):
    """gatconv_layer_fused.py:89-118; params = preprocess_Hyper(g)."""

    def conv(self, indptr, indices, a_l, a_r, h):
        attn_row, attn_col = self._attn(a_l, a_r, h)
        return GATConvFuse_inference_hyper_recompute(attn_row, attn_col, indptr, indices,
                                                     self.negative_slope, h)

    def forward(self, params, feat, fuse=False):
        args = params[:2] if fuse else ()
        return self._forward(self.conv, args, params, feat, fuse)


class GATConv_softmax(GATConvDGL):
    """gatconv_layer_fused.py:120-156; params = preprocess_softmax(g)."""

    def conv(self, indptr, indices, rows, smem_consume, a_l, a_r, h):
        attn_row, attn_col = self._attn(a_l, a_r, h)
        return GATConvFuse_inference_softmax(smem_consume, attn_row, attn_col, indptr, indices,
                                             rows, self.negative_slope, h)

    def forward(self, params, feat, fuse=False):
        args = (params[0], params[1], params[2], params[4]) if fuse else ()
        return self._forward(self.conv, args, params, feat, fuse)


class GATConv_softmax_gm(GATConvDGL):
    """gatconv_layer_softmax_gm.py:7-41; params = preprocess_softmax(g)."""

    def conv(self, indptr, indices, rows, a_l, a_r, h):
        attn_row, attn_col = self._attn(a_l, a_r, h)
        return GATConvFuse_inference_softmax_gm(attn_row, attn_col, indptr, indices, rows,
                                                self.negative_slope, h)

    def forward(self, params, feat, fuse=False):
        args = params[:3] if fuse else ()
        return self._forward(self.conv, args, params, feat, fuse)


class GATConv_hyper_v2(GATConvDGL):
    """gatconv_layer_fused.py:159-191: attention logits computed inside the op."""

    def conv(self, indptr, indices, smem_consume, a_l, a_r, h):
        return GATConvFuse_inference_hyper_v2(smem_consume, a_l.contiguous(), a_r.contiguous(),
                                              indptr, indices, self.negative_slope, h)

    def forward(self, params, feat, fuse=False):
        args = (params[0], params[1], params[4]) if fuse else ()
        return self._forward(self.conv, args, params, feat, fuse)


class GATConv_forward(GATConvDGL):
    """Training module over FusedGATFunction (the layer the reference's
    script/train/train_gatconv.py:10 imports but does not ship).
    params = preprocess_gat_fw_bw(g) = (row_ptr, col_ind, col_ptr, row_ind, permute).
    ``fused_projection = True``: feat = W x and both logit vectors come from ONE tensor-core kernel
    (operators/projection.py; the GEMM epilogue reduces <a_l, feat>, <a_r, feat> from the accumulator
    rows) instead of a cuBLAS GEMM + two elementwise-reduce kernels."""

    fused_projection = False

    def __init__(self, in_size, out_size, num_heads, dropout=0, negative_slope=0.2):
        super().__init__(in_size, out_size, num_heads, dropout, negative_slope)
        self.attn_drop = float(dropout)

    def _project(self, feat):
        n_out = self.out_size * self.num_heads
        hd = self.out_size
        if (self.fused_projection and feat.is_cuda and projection.supported(self.in_size, n_out, 1)
                and (hd in (8, 16) or hd % 32 == 0)):
            cache = self.__dict__.setdefault("_proj_cache", projection.PackedWeights())
            return projection.FusedGATProjFunction.apply(
                feat, self.W.weight, self.W.bias, self.a_l.transpose(1, 2), self.a_r.transpose(1, 2),
                self.num_heads, cache)
        h = self.W(feat).view(-1, self.num_heads, self.out_size).contiguous()
        return h, (self.a_l.transpose(1, 2) * h).sum(dim=-1), (self.a_r.transpose(1, 2) * h).sum(dim=-1)

    def forward(self, params, feat, fuse=True):
        N = len(feat)
        row_ptr, col_ind, col_ptr, row_ind, permute = params
        h, attn_row, attn_col = self._project(feat)
        drop = self.attn_drop if self.training else 0.0
        out = GATConvFuse(attn_row.contiguous(), attn_col.contiguous(), row_ptr, col_ind, col_ptr,
                          row_ind, permute, self.negative_slope, h, drop)
        return out.reshape(N, -1)
