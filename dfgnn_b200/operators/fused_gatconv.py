"""Drop-in for ``DFGNN/operators/fused_gatconv.py``: same names, argument order and
returns.  ``fused_gat`` stands for the reference's pybind module ``fused_gatconv``."""
import torch

from . import _native as fused_gat


def GATConvFuse(attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind, permute, negative_slope,
                in_feat, attn_drop):
    """operators/fused_gatconv.py:5-28."""
    return FusedGATFunction.apply(attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind, permute,
                                  negative_slope, in_feat, attn_drop)


def GATConvFuse_inference_hyper(smem_consume, attn_row, attn_col, indptr, indices, rows,
                                negative_slope, in_feat):
    """operators/fused_gatconv.py:31-36."""
    return fused_gat.gat_inference_hyper(smem_consume, attn_row, attn_col, indptr, indices, rows,
                                         negative_slope, in_feat)


def GATConvFuse_inference_hyper_recompute(attn_row, attn_col, indptr, indices, negative_slope,
                                          in_feat):
    """operators/fused_gatconv.py:39-44."""
    return fused_gat.gat_inference_hyper_recompute(attn_row, attn_col, indptr, indices,
                                                   negative_slope, in_feat)


def GATConvFuse_inference_hyper_v2(smem_consume, a_l, a_r, indptr, indices, negative_slope,
                                   in_feat):
    """operators/fused_gatconv.py:47-52."""
    return fused_gat.gat_inference_hyper_v2(smem_consume, a_l, a_r, indptr, indices,
                                            negative_slope, in_feat)


def GATConvFuse_inference_softmax(smem_consume, attn_row, attn_col, indptr, indices, rows,
                                  negative_slope, in_feat):
    """operators/fused_gatconv.py:63-68."""
    return fused_gat.gat_inference_softmax(smem_consume, attn_row, attn_col, indptr, indices, rows,
                                           negative_slope, in_feat)


def GATConvFuse_inference_softmax_gm(attn_row, attn_col, indptr, indices, rows, negative_slope,
                                     in_feat):
    """operators/fused_gatconv.py:71-76."""
    return fused_gat.gat_inference_softmax_gm(attn_row, attn_col, indptr, indices, rows,
                                              negative_slope, in_feat)


def GATConvFuse_inference_tiling(attn_row, attn_col, row_ptr, col_ind, negative_slope, in_feat):
    """operators/fused_gatconv.py:79-84."""
    return fused_gat.gat_inference_tiling(attn_row, attn_col, row_ptr, col_ind, negative_slope,
                                          in_feat)


def GATConvFuse_inference(attn_row, attn_col, row_ptr, col_ind, negative_slope, in_feat):
    """operators/fused_gatconv.py:87-92."""
    return fused_gat.gat_inference(attn_row, attn_col, row_ptr, col_ind, negative_slope, in_feat)


class FusedGATFunction(torch.autograd.Function):
    """operators/fused_gatconv.py:95-176."""

    @staticmethod
    def forward(ctx, attn_row, attn_col, row_ptr, col_ind, col_ptr, row_ind, permute,
                negative_slope, in_feat, attn_drop):
        out_feat, edge_max, edge_sum, edge_mask = fused_gat.gat_forward(
            attn_row, attn_col, row_ptr, col_ind, negative_slope, in_feat, attn_drop)
        ctx.save_for_backward(row_ptr, col_ind, col_ptr, row_ind, permute, edge_max, edge_sum,
                              edge_mask, in_feat, attn_row, attn_col)
        ctx.negative_slope = negative_slope
        ctx.attn_drop = attn_drop
        return out_feat

    @staticmethod
    def backward(ctx, grad_out):
        (row_ptr, col_ind, col_ptr, row_ind, permute, edge_max, edge_sum, edge_mask, in_feat,
         attn_row, attn_col) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_feat, grad_attn_row, grad_attn_col = fused_gat.gat_backward(
            ctx.negative_slope, ctx.attn_drop, row_ptr, col_ind, col_ptr, row_ind, permute,
            edge_max, edge_sum, edge_mask, in_feat, attn_row, attn_col, grad_out)
        return (grad_attn_row, grad_attn_col, None, None, None, None, None, None, grad_feat, None)
