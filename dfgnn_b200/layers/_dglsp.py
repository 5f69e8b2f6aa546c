"""Torch restatement of the three dgl.sparse calls of the reference's NON-fused
branch (``bsddmm`` -> ``softmax`` -> ``bspmm``; ``layers/GT/gtconv_layer.py:29-33``).
It exists so that the conv modules keep their ``fuse=False`` comparison branch
(the reference checks fused against it with ``check_correct``); it is not a
fallback for the fused operators and is never called by them."""
from __future__ import annotations

import torch


def edge_softmax(A, score: torch.Tensor) -> torch.Tensor:
    """Row-wise softmax of per-edge scores [E, nh] (dgl.sparse ``.softmax()``)."""
    n = A.shape[0]
    row = A.row.long()
    idx = row.unsqueeze(1).expand_as(score)
    mx = torch.full((n, score.shape[1]), float("-inf"), dtype=score.dtype, device=score.device)
    mx = mx.scatter_reduce(0, idx, score, reduce="amax", include_self=True)
    ex = torch.exp(score - mx[row])
    sm = torch.zeros_like(mx).index_add_(0, row, ex)
    return ex / sm[row]


def bsddmm(A, q: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """Per-edge <q[row], k[col]> per head; q, k are [N, d, nh]."""
    return (q[A.row.long()] * k[A.col.long()]).sum(dim=1)


def bspmm(A, attn: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """out[row] += attn[e] * v[col]; v is [N, d, nh], attn [E, nh]."""
    out = torch.zeros_like(v)
    return out.index_add_(0, A.row.long(), attn.unsqueeze(1) * v[A.col.long()])
