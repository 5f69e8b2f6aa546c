"""Multi-GPU partitioning of the conv (one process per GPU, torch.distributed / NCCL).

The reference is single-GPU (no distributed code at all), so this
is new design, following SURVEY.md 8(e):

* **Batched datasets** are block diagonal: whole graphs go to ranks, every rank
  builds its own CSR/CSC and runs the conv with **no collective**.
* **Full graphs** are 1-D **row partitioned**, boundaries chosen on the degree
  prefix sum so that every rank holds ~E/P edges.  Rank r owns rows [lo_r, hi_r):
  Q / out / dO for those rows and the slice of the column-side operands (K, V or
  feat, attn_col) of the same nodes.  Forward: an all-gather of the column-side
  operands ("halo"; for the dense-halo graphs of the benchmark the halo is every
  node).  Backward: the column-indexed partial gradients are reduce-scattered.
  Owned slices have different lengths, so they are padded to ``max_rows`` and the
  shard's column ids are relabelled once to ``owner * max_rows + local`` -- the
  gathered buffer is then indexed directly by the kernels, no unpacking pass.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .graphs import Graph


@dataclass
class Partition:
    rank: int
    world: int
    kind: str                  # "single" | "by-graph" | "row"
    scaling: str               # "weak" | "strong"
    local_graph: Graph         # rows local, columns in the (padded) gathered index space
    n_rows: int                # local rows
    n_cols: int                # columns of the local matrix
    row_slice: slice           # this rank's rows in the global node arrays
    col_owned: slice           # this rank's slice of the column-side operands
    max_rows: int              # padded slice length (row partition)
    bounds: Optional[torch.Tensor]  # [world+1] row boundaries (row partition)
    describe: str


def row_bounds(deg: torch.Tensor, world: int) -> torch.Tensor:
    """nnz-balanced contiguous row boundaries: bounds[r] = first row of rank r."""
    csum = torch.cumsum(deg.to(torch.int64), 0)
    total = int(csum[-1]) if csum.numel() else 0
    targets = torch.arange(1, world, dtype=torch.int64) * total // world
    cuts = torch.searchsorted(csum, targets, right=False) + 1 if world > 1 else targets
    b = torch.cat([torch.zeros(1, dtype=torch.int64), cuts.clamp(max=deg.numel()),
                   torch.tensor([deg.numel()], dtype=torch.int64)])
    return torch.cummax(b, 0).values


def graph_bounds(g: Graph, world: int) -> torch.Tensor:
    """Contiguous split of the graphs of a batch, balanced by edges: -> [world+1] graph ids."""
    bnn = g.batch_num_nodes().to(torch.int64).cpu()
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(bnn, 0)])
    src = g.edges()[0].cpu()
    node_deg = torch.bincount(src, minlength=g.num_nodes())
    gid = torch.repeat_interleave(torch.arange(bnn.numel()), bnn)
    edges_per_graph = torch.zeros(bnn.numel(), dtype=torch.int64).index_add_(0, gid, node_deg)
    return row_bounds(edges_per_graph, world), offs


def make_partition(g: Graph, world: int, rank: int, mode: str = "auto") -> Partition:
    """Shard ``g`` (a CPU graph with canonical edge order) for ``rank`` of ``world``."""
    n = g.num_nodes()
    if world == 1:
        return Partition(rank, world, "single", "weak", g, n, g.num_cols, slice(0, n), slice(0, n),
                         n, None, "single GPU, whole graph")
    src, dst = (t.cpu() for t in g.edges())
    if mode == "auto":
        mode = "by-graph" if g.batch_size > 1 else "row"
    if mode == "by-graph":
        gb, offs = graph_bounds(g, world)
        g_lo, g_hi = int(gb[rank]), int(gb[rank + 1])
        lo, hi = int(offs[g_lo]), int(offs[g_hi])
        keep = (src >= lo) & (src < hi)
        bnn = g.batch_num_nodes()[g_lo:g_hi]
        local = Graph(src[keep] - lo, dst[keep] - lo, hi - lo, bnn, g.name + f"[graphs {g_lo}:{g_hi}]")
        return Partition(rank, world, "by-graph", "strong", local, hi - lo, hi - lo, slice(lo, hi),
                         slice(lo, hi), hi - lo, None,
                         f"global batch sharded by whole graph over {world} ranks (edge balanced), "
                         f"no collective")
    deg = torch.bincount(src, minlength=n)
    bounds = row_bounds(deg, world)
    sizes = bounds[1:] - bounds[:-1]
    max_rows = int(sizes.max())
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    keep = (src >= lo) & (src < hi)
    d = dst[keep]
    owner = torch.searchsorted(bounds[1:].contiguous(), d, right=True)
    col = owner * max_rows + (d - bounds[owner])
    local = Graph(src[keep] - lo, col, hi - lo, None, g.name + f"[rows {lo}:{hi}]",
                  num_cols=world * max_rows)
    return Partition(rank, world, "row", "strong", local, hi - lo, world * max_rows, slice(lo, hi),
                     slice(lo, hi), max_rows, bounds,
                     f"1-D row partition over {world} ranks (nnz balanced), halo all-gather of the "
                     f"column-side operands + reduce-scatter of their gradients")


class HaloExchange:
    """All-gather of the column-side operands / reduce-scatter of their gradients for a
    row partition.  Identity for the other partition kinds.  Works on any backend
    (NCCL on the GPUs; gloo in the CPU tests, where reduce-scatter is an all-reduce + slice)."""

    def __init__(self, part: Partition, device, world: int):
        self.part = part
        self.active = part.kind == "row" and world > 1
        self.device = device
        self.world = world
        self._send = {}
        self._recv = {}

    def _buf(self, store, key, shape, like):
        t = store.get(key)
        if t is None or t.shape != torch.Size(shape) or t.dtype != like.dtype:
            t = torch.zeros(shape, dtype=like.dtype, device=like.device)
            store[key] = t
        return t

    def gather(self, x: torch.Tensor, key: str = "a") -> torch.Tensor:
        """x: this rank's owned slice [n_owned, ...] -> [world*max_rows, ...] in padded order."""
        import torch.distributed as dist
        if not self.active:
            return x
        mr = self.part.max_rows
        send = self._buf(self._send, key, (mr,) + tuple(x.shape[1:]), x)
        send[: x.shape[0]].copy_(x)
        out = self._buf(self._recv, key, (self.world * mr,) + tuple(x.shape[1:]), x)
        dist.all_gather_into_tensor(out, send)
        return out

    def gather_pair(self, a, b, rec=None):
        if not self.active:
            if rec is not None and "ag0" in rec:
                rec["ag0"].record()
                rec["ag1"].record()
            return a, b
        if rec is not None:
            rec["ag0"].record()
        oa, ob = self.gather(a, "a"), self.gather(b, "b")
        if rec is not None:
            rec["ag1"].record()
        return oa, ob

    def reduce(self, g: torch.Tensor, key: str = "ga") -> torch.Tensor:
        """g: partial gradient over ALL padded columns [world*max_rows, ...] -> owned slice."""
        import torch.distributed as dist
        if not self.active:
            return g
        mr = self.part.max_rows
        n_owned = self.part.n_rows
        if dist.get_backend() == "gloo":
            dist.all_reduce(g)
            r = self.part.rank
            return g[r * mr: r * mr + n_owned]
        out = self._buf(self._recv, key, (mr,) + tuple(g.shape[1:]), g)
        dist.reduce_scatter_tensor(out, g.contiguous())
        return out[:n_owned]

    def reduce_pair(self, ga, gb, rec=None):
        if not self.active:
            if rec is not None and "rs0" in rec:
                rec["rs0"].record()
                rec["rs1"].record()
            return ga, gb
        if rec is not None:
            rec["rs0"].record()
        oa, ob = self.reduce(ga, "ga"), self.reduce(gb, "gb")
        if rec is not None:
            rec["rs1"].record()
        return oa, ob
