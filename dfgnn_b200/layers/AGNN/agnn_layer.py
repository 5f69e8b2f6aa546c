"""AGNN conv modules: names and contract of ``DFGNN/layers/AGNN/*.py``.

The reference normalises H with ``F.normalize`` and feeds (H_norm, H_norm, H) to
the GT operators; here the fused branches call one kernel that normalises on
the fly and reuses each gathered row of H for both the score and the
aggregation (``AGNNConvFuse_inference``).  Set ``fuse_normalize=False`` on a
module to get the reference's literal two-step formulation through the GT
operator of the same format."""
from torch import nn
from torch.nn import functional as F

from ...operators.fused_gtconv import (AGNNConvFuse_inference, GTConvFuse_hyper,
                                       GTConvFuse_inference_csr, GTConvFuse_inference_csr_gm,
                                       GTConvFuse_inference_hyper, GTConvFuse_inference_softmax,
                                       GTConvFuse_inference_softmax_gm,
                                       GTConvFuse_inference_tiling)
from ...utils import benchmark
from .._dglsp import bsddmm, bspmm, edge_softmax


class AGNNConvDGL(nn.Module):
    """layers/AGNN/agnn_layer.py:6-19."""

    fuse_normalize = True

    def __init__(self, in_size, out_size, num_heads):
        super().__init__()
        self.in_size = in_size
        self.out_size = out_size
        self.num_heads = num_heads
        self.proj = nn.Linear(in_size, out_size)

    def forward_dglsp(self, A, H):
        """agnn_layer.py:14-19; H is [N, out, heads]."""
        H_norm = F.normalize(H, p=2, dim=1)
        attn = edge_softmax(A, bsddmm(A, H_norm, H_norm))
        return bspmm(A, attn, H)

    # how the module's format maps params -> (indptr, indices) and the GT operator call
    def _gt_call(self, params, H_norm, H):
        raise NotImplementedError

    def conv(self, H, *params):
        """agnn_layer_fused.py:13-25."""
        if self.fuse_normalize:
            return AGNNConvFuse_inference(params[0], params[1], H)
        return self._gt_call(params, F.normalize(H, p=2, dim=-1), H)

    def forward(self, params, feat, fuse=False):
        """agnn_layer_fused.py:27-46."""
        N = len(feat)
        H = self.proj(feat).view(-1, self.num_heads, self.out_size)
        if fuse:
            H = H.contiguous()
            out, elapsed_time = benchmark(self.conv, H, *params)
        else:
            H = H.reshape(-1, self.out_size, self.num_heads)
            out, elapsed_time = benchmark(self.forward_dglsp, params, H)
            out = out.transpose(1, 2)
        return out.reshape(N, -1), elapsed_time * 1000


class AGNNConv_csr(AGNNConvDGL):
    """agnn_layer_fused.py:12-46; params = preprocess_CSR(g)."""

    def _gt_call(self, params, Hn, H):
        indptr, indices, val, smem_consume = params
        return GTConvFuse_inference_csr(indptr, indices, val, smem_consume, Hn, Hn, H)


class AGNNConv_csr_gm(AGNNConvDGL):
    """agnn_layer_csr_gm.py; params = preprocess_CSR(g)."""

    def _gt_call(self, params, Hn, H):
        indptr, indices, val, _ = params
        return GTConvFuse_inference_csr_gm(indptr, indices, val, Hn, Hn, H)


class AGNNConv_tiling(AGNNConvDGL):
    """agnn_layer_tiling.py:9-42; params = preprocess_CSR(g)."""

    def _gt_call(self, params, Hn, H):
        indptr, indices, val, smem_consume = params
        return GTConvFuse_inference_tiling(indptr, indices, val, smem_consume, Hn, Hn, H)


class AGNNConv_softmax(AGNNConvDGL):
    """agnn_layer_fused.py:49-85; params = preprocess_softmax(g)."""

    def _gt_call(self, params, Hn, H):
        indptr, indices, rows, val, smem_consume = params
        return GTConvFuse_inference_softmax(indptr, indices, rows, val, smem_consume, Hn, Hn, H)


class AGNNConv_softmax_gm(AGNNConvDGL):
    """agnn_layer_softmax_gm.py; params = preprocess_softmax(g)."""

    def _gt_call(self, params, Hn, H):
        indptr, indices, rows, val, _ = params
        return GTConvFuse_inference_softmax_gm(indptr, indices, rows, val, Hn, Hn, H)


class AGNNConv_hyper(AGNNConvDGL):
    """agnn_layer_fused.py:88-121; params = preprocess_Hyper(g)."""

    def _gt_call(self, params, Hn, H):
        indptr, indices, rows, val, smem_consume = params
        return GTConvFuse_inference_hyper(indptr, indices, rows, val, smem_consume, Hn, Hn, H)


class AGNNConv_forward(AGNNConvDGL):
    """Training module (agnn_layer_forward.py:8-66); params = preprocess_Hyper_fw_bw(g).
    Gradients flow through F.normalize by autograd, as in the reference."""

    def conv(self, H, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem_consume):
        H_norm = F.normalize(H, p=2, dim=-1)
        return GTConvFuse_hyper(rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx,
                                smem_consume, H_norm, H_norm, H)

    def forward(self, params, feat, fuse=False):
        N = len(feat)
        A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem_consume = params
        if fuse:
            H = self.proj(feat).view(-1, self.num_heads, self.out_size)
            out = self.conv(H, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem_consume)
        else:
            H = self.proj(feat).view(-1, self.out_size, self.num_heads)
            out = self.forward_dglsp(A, H)
        return out.reshape(N, -1)
