"""Drop-in for ``DFGNN/operators/fused_gtconv.py``: same function names, argument
order and return values (reference file:line in each docstring).  ``fused_gt``
below plays the role of the reference's pybind module ``fused_gtconv``."""
import torch

from . import _native as fused_gt


def GTConvFuse_inference_hyper(indptr, indices, rows, val, smem_consume, Q, K, V):
    """operators/fused_gtconv.py:5-25."""
    return fused_gt.gt_hyper_inference(indptr, indices, rows, val, smem_consume, Q, K, V)[0]


def GTConvFuse_hyper(rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem_consume, Q, K, V):
    """operators/fused_gtconv.py:51-76 (note: `rows` comes FIRST here, third natively)."""
    return FusedGTFunction_hyper.apply(rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx,
                                       smem_consume, Q, K, V)


class FusedGTFunction_hyper(torch.autograd.Function):
    """operators/fused_gtconv.py:79-158."""

    @staticmethod
    def forward(ctx, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem_consume, Q, K, V):
        out_feat, attn_edge = fused_gt.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr,
                                                        row_ind, val_idx, smem_consume, Q, K, V)
        ctx.smem = smem_consume
        ctx.save_for_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, Q, K, V,
                              attn_edge)
        return out_feat

    @staticmethod
    def backward(ctx, grad_out):
        (row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, Q, K, V,
         attn_edge) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_Q, grad_K, grad_V = fused_gt.gt_backward(row_ptr, col_ind, rows, val, col_ptr,
                                                      row_ind, val_idx, ctx.smem, Q, K, V,
                                                      attn_edge, grad_out)
        return (None, None, None, None, None, None, None, None, grad_Q, grad_K, grad_V)


def GTConvFuse_inference_softmax(indptr, indices, rows, val, smem_consume, Q, K, V):
    """operators/fused_gtconv.py:238-258."""
    return fused_gt.gt_softmax_inference(indptr, indices, rows, val, smem_consume, Q, K, V)[0]


def GTConvFuse_inference_softmax_gm(indptr, indices, rows, val, Q, K, V):
    """operators/fused_gtconv.py:261-279."""
    return fused_gt.gt_softmax_gm_inference(indptr, indices, rows, val, Q, K, V)


def GTConvFuse_inference_csr(indptr, indices, val, smem_consume, Q, K, V):
    """operators/fused_gtconv.py:282-301."""
    return fused_gt.gt_csr_inference(indptr, indices, val, smem_consume, Q, K, V)[0]


def GTConvFuse_inference_csr_gm(indptr, indices, val, Q, K, V):
    """operators/fused_gtconv.py:304-321."""
    return fused_gt.gt_csr_gm_inference(indptr, indices, val, Q, K, V)[0]


def GTConvFuse_inference_tiling(indptr, indices, val, smem_consume, Q, K, V):
    """operators/fused_gtconv.py:324-343."""
    return fused_gt.gt_tiling_inference(indptr, indices, val, smem_consume, Q, K, V)[0]


def AGNNConvFuse_inference(indptr, indices, H):
    """AGNN conv with the L2 normalisation fused in (replaces F.normalize + the GT
    call of layers/AGNN/agnn_layer_fused.py:13-25; no reference operator)."""
    return fused_gt.agnn_forward(indptr, indices, H)
