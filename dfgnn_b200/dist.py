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

Padded column space.  Owned slices have different lengths, so every rank keeps its
slice in a buffer of ``max_rows`` rows (zero tail) and the shard's column ids are
relabelled ONCE so that the gathered buffer is indexed directly by the kernels (no
unpacking pass).  The padded space is cut into ``chunks`` equal column chunks:

    local row l of owner w  ->  c = l // q, lq = l % q          (q = max_rows / chunks)
    column id               =   c * (world * q) + w * q + lq

i.e. chunk c of the gathered buffer is the all-gather of every rank's c-th slice
``x[c*q:(c+1)*q]`` -- one ``ncclAllGather`` per chunk and operand, all of them issued as
ONE coalesced NCCL group -- and chunk c of the column-side gradients is reduce-scattered
on a side stream while the column-side kernel of chunk c+1 runs (``chunks`` > 1).

Public operators (autograd Functions, same argument meaning as the single-GPU
``GTConvFuse_hyper`` / ``GATConvFuse`` with the column-side operands given as owned slices):
``GTConvFuse_hyper_dist`` and ``GATConvFuse_dist``; ``HaloExchange.gather`` on its own is
differentiable too (its backward is the reduce-scatter).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

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
    chunks: int = 1            # column chunks of the padded space (row partition)

    @property
    def q(self) -> int:
        """rows of one chunk slice"""
        return self.max_rows // self.chunks

    def padded_index(self, node: torch.Tensor) -> torch.Tensor:
        """global node id -> column id in the padded gathered space (row partition)."""
        b = self.bounds
        owner = torch.searchsorted(b[1:].contiguous(), node, right=True)
        loc = node - b[owner]
        q = self.q
        return (loc // q) * (self.world * q) + owner * q + (loc % q)


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


def make_partition(g: Graph, world: int, rank: int, mode: str = "auto", chunks: int = 1) -> Partition:
    """Shard ``g`` (a CPU graph with canonical edge order) for ``rank`` of ``world``.
    ``chunks``: column chunks of the padded space of a row partition (see module docstring)."""
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
    chunks = max(1, int(chunks))
    deg = torch.bincount(src, minlength=n)
    bounds = row_bounds(deg, world)
    sizes = bounds[1:] - bounds[:-1]
    max_rows = (int(sizes.max()) + chunks - 1) // chunks * chunks
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    keep = (src >= lo) & (src < hi)
    part = Partition(rank, world, "row", "strong", g, hi - lo, world * max_rows, slice(lo, hi),
                     slice(lo, hi), max_rows, bounds,
                     f"1-D row partition over {world} ranks (nnz balanced), halo all-gather of the "
                     f"column-side operands + reduce-scatter of their gradients"
                     + (f", {chunks} column chunks (reduce-scatter overlapped with the column-side kernels)"
                        if chunks > 1 else ""), chunks)
    col = part.padded_index(dst[keep])
    part.local_graph = Graph(src[keep] - lo, col, hi - lo, None, g.name + f"[rows {lo}:{hi}]",
                             num_cols=world * max_rows)
    return part


# --------------------------------------------------------------------------------------- #
# halo exchange                                                                            #
# --------------------------------------------------------------------------------------- #

class HaloExchange:
    """All-gather of the column-side operands / reduce-scatter of their gradients for a
    row partition.  Identity for the other partition kinds.  Works on any backend
    (NCCL on the GPUs; gloo in the CPU tests, where the collectives are emulated with
    all_gather / all_reduce + slice).

    Buffers: every call returns FRESH tensors (they may be saved for backward by the conv
    Functions); an operand that already has ``max_rows`` rows is sent in place, a shorter one
    (the ``n_rows`` owned rows) is zero-padded first.

    ``record=True`` brackets every collective with CUDA events on the stream it runs on;
    ``pop_times()`` returns the accumulated {"allgather_ms", "reduce_scatter_ms"}."""

    def __init__(self, part: Partition, device, world: int, record: bool = False, backend: str = "auto"):
        import os

        import torch.distributed as dist
        self.part = part
        self.active = part.kind == "row" and world > 1
        self.device = torch.device(device)
        self.world = world
        self.record = record and self.device.type == "cuda"
        self._events = {"allgather_ms": [], "reduce_scatter_ms": []}
        self._nccl = self.active and dist.get_backend() == "nccl"
        self._comm = torch.cuda.Stream(self.device) if (self.active and self.device.type == "cuda") else None
        # "nccl": the coalesced NCCL collectives.  "p2p": pull-based exchange over NVLink peer memory
        # (torch symmetric memory: every rank maps the others' buffers and copies its blocks out of
        # them, no NCCL launch).  Measured on 8 B200, reddit-shaped graph: NCCL
        # 5.59 ms per step, p2p 5.73 ms -- the peer READS are latency bound -- so "auto" = nccl and p2p
        # stays an option (DFGNN_B200_HALO=p2p, or backend="p2p").
        backend = os.environ.get("DFGNN_B200_HALO", backend)
        if backend == "auto":
            backend = "nccl"
        self.backend = "nccl"
        self.backend_note = None
        self._sym = {}
        if self._nccl and backend in ("auto", "p2p") and part.chunks == 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._symm_mem = symm_mem
                probe = symm_mem.empty((1024,), dtype=torch.float32, device=self.device)
                symm_mem.rendezvous(probe, dist.group.WORLD.group_name).barrier()
                self.backend = "p2p"
            except Exception as exc:  # no peer mapping on this system: NCCL
                self.backend_note = "symmetric memory unavailable (%s)" % repr(exc)[:120]
                if backend == "p2p":
                    raise

    # -- peer-memory exchange (backend "p2p") -----------------------------------------------------
    def _sym_buf(self, key, shape, dtype):
        """A persistent symmetric buffer (same shape on every rank) and its rendezvous handle."""
        import torch.distributed as dist
        ent = self._sym.get(key)
        if ent is None or ent[0].shape != torch.Size(shape) or ent[0].dtype != dtype:
            t = self._symm_mem.empty(tuple(shape), dtype=dtype, device=self.device)
            ent = (t, self._symm_mem.rendezvous(t, dist.group.WORLD.group_name))
            self._sym[key] = ent
        return ent

    def _p2p_all_gather(self, xs):
        mr, W, r = self.part.max_rows, self.world, self.part.rank
        outs = []
        hdls = []
        for i, x in enumerate(xs):  # own slice -> symmetric memory (one local copy)
            buf, hdl = self._sym_buf(("ag", i), x.shape, x.dtype)
            buf.copy_(x)
            hdls.append((buf, hdl))
            outs.append(x.new_empty((W * mr,) + tuple(x.shape[1:])))
        hdls[0][1].barrier()  # every rank's slices are in place
        for step in range(W):  # pull, starting with the own slice, every rank from a different peer
            peer = (r - step) % W
            for (buf, hdl), x, o in zip(hdls, xs, outs):
                src = buf if peer == r else hdl.get_buffer(peer, x.shape, x.dtype)
                o[peer * mr:(peer + 1) * mr].copy_(src)
        hdls[0][1].barrier()  # nobody overwrites a slice that is still being read
        return outs

    def p2p_grad_buffers(self, likes):
        """Persistent symmetric buffers for the column-side partial gradients ([world*max_rows, ...]
        each): the column-side kernels write them in place and the peers pull their blocks."""
        return [self._sym_buf(("rs", i), (self.world * self.part.max_rows,) + tuple(t.shape[1:]), t.dtype)[0]
                for i, t in enumerate(likes)]

    def _p2p_reduce_scatter(self, gs):
        mr, W, r = self.part.max_rows, self.world, self.part.rank
        ents = []
        for i, g in enumerate(gs):
            buf, hdl = self._sym_buf(("rs", i), g.shape, g.dtype)
            if g.data_ptr() != buf.data_ptr():
                buf.copy_(g)  # a caller that did not write into p2p_grad_buffers(): one local copy
            ents.append((buf, hdl))
        ents[0][1].barrier()
        outs = []
        stage = [g.new_empty((W, mr) + tuple(g.shape[1:])) for g in gs]
        for step in range(W):
            peer = (r - step) % W
            for (buf, hdl), g, st in zip(ents, gs, stage):
                blk = (mr,) + tuple(g.shape[1:])
                n_blk = mr * math.prod(g.shape[1:])
                src = buf[r * mr:(r + 1) * mr] if peer == r else hdl.get_buffer(peer, blk, g.dtype, n_blk * r)
                st[peer].copy_(src)
        ents[0][1].barrier()
        for st in stage:
            outs.append(st.sum(dim=0))
        return outs

    # -- helpers ---------------------------------------------------------------------------
    def chunk_nnz(self, col_ptr: torch.Tensor, c: int) -> int:
        """Entries of column chunk c of the shard's CSC (schedule heuristics of the column-side
        launches); read back once per col_ptr tensor and cached."""
        key = (col_ptr.data_ptr(), col_ptr.numel())
        if getattr(self, "_nnz_key", None) != key:
            w = self.world * self.part.q
            cuts = col_ptr[:: w].tolist() if self.part.chunks > 1 else [0, int(col_ptr[-1])]
            if len(cuts) < self.part.chunks + 1:
                cuts.append(int(col_ptr[-1]))
            self._nnz_key, self._nnz = key, [cuts[i + 1] - cuts[i] for i in range(self.part.chunks)]
        return self._nnz[c]

    def pad(self, x: torch.Tensor) -> torch.Tensor:
        """owned slice [n_rows, ...] -> [max_rows, ...] (zero tail); already padded: unchanged."""
        mr = self.part.max_rows
        if not self.active or x.shape[0] == mr:
            return x
        if x.shape[0] != self.part.n_rows:
            raise RuntimeError(f"halo operand has {x.shape[0]} rows, expected {self.part.n_rows} or {mr}")
        out = x.new_zeros((mr,) + tuple(x.shape[1:]))
        out[: x.shape[0]].copy_(x)
        return out

    def _ev(self, key):
        if not self.record:
            return None
        pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._events[key].append(pair)
        pair[0].record()
        return pair

    def pop_times(self) -> dict:
        """Sum of the recorded collective durations since the last call (synchronises)."""
        out = {}
        if self.record:
            torch.cuda.synchronize(self.device)
        for k, pairs in self._events.items():
            out[k] = float(sum(a.elapsed_time(b) for a, b in pairs))
            pairs.clear()
        return out

    # -- raw collectives ---------------------------------------------------------------------
    def all_gather(self, xs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """xs: owned slices -> gathered [world*max_rows, ...] tensors in the padded order."""
        import torch.distributed as dist
        if not self.active:
            return list(xs)
        C, q, W = self.part.chunks, self.part.q, self.world
        xs = [self.pad(x.detach()).contiguous() for x in xs]
        if self.backend == "p2p":
            pair = self._ev("allgather_ms")
            outs = self._p2p_all_gather(xs)
            if pair:
                pair[1].record()
            return outs
        outs = [x.new_empty((C, W * q) + tuple(x.shape[1:])) for x in xs]
        pair = self._ev("allgather_ms")
        if self._nccl:
            with dist._coalescing_manager():  # one NCCL group launch for all chunks and operands
                for x, o in zip(xs, outs):
                    for c in range(C):
                        dist.all_gather_into_tensor(o[c], x[c * q:(c + 1) * q])
        else:
            for x, o in zip(xs, outs):
                for c in range(C):
                    parts = [torch.empty_like(x[:q]) for _ in range(W)]
                    dist.all_gather(parts, x[c * q:(c + 1) * q].contiguous())
                    o[c].copy_(torch.cat(parts, 0))
        if pair:
            pair[1].record()
        return [o.view((W * C * q,) + tuple(o.shape[2:])) for o in outs]

    def reduce_scatter_chunk(self, c: int, gs: Sequence[torch.Tensor], outs: Sequence[torch.Tensor]):
        """Chunk c of the column-indexed partial gradients gs ([world*max_rows, ...]) summed over
        the ranks into rows [c*q, (c+1)*q) of outs ([max_rows, ...]).  Runs on the current stream."""
        import torch.distributed as dist
        C, q, W = self.part.chunks, self.part.q, self.world
        pair = self._ev("reduce_scatter_ms")
        if self._nccl:
            with dist._coalescing_manager():
                for g, o in zip(gs, outs):
                    dist.reduce_scatter_tensor(o[c * q:(c + 1) * q], g[c * W * q:(c + 1) * W * q])
        else:
            r = self.part.rank
            for g, o in zip(gs, outs):
                blk = g[c * W * q:(c + 1) * W * q].clone()
                dist.all_reduce(blk)
                o[c * q:(c + 1) * q].copy_(blk[r * q:(r + 1) * q])
        if pair:
            pair[1].record()

    def reduce_scatter(self, gs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """gs: partial gradients over ALL padded columns -> owned slices [n_rows, ...]."""
        if not self.active:
            return list(gs)
        mr = self.part.max_rows
        gs = [g.contiguous() for g in gs]
        if self.backend == "p2p":
            pair = self._ev("reduce_scatter_ms")
            outs = self._p2p_reduce_scatter(gs)
            if pair:
                pair[1].record()
            return [o[: self.part.n_rows] for o in outs]
        outs = [g.new_empty((mr,) + tuple(g.shape[1:])) for g in gs]
        for c in range(self.part.chunks):
            self.reduce_scatter_chunk(c, gs, outs)
        return [o[: self.part.n_rows] for o in outs]

    # -- overlapped reduce-scatter (used by the distributed conv Functions) --------------------
    def begin_overlapped_reduce(self, gs: Sequence[torch.Tensor]):
        mr = self.part.max_rows
        self._ov = ([g.new_empty((mr,) + tuple(g.shape[1:])) for g in gs], list(gs))
        return self._ov[0]

    def reduce_chunk_async(self, c: int):
        """Called after the kernels that produce chunk c were launched on the current stream:
        its reduce-scatter runs on the communication stream behind them."""
        outs, gs = self._ov
        if self._comm is None:
            self.reduce_scatter_chunk(c, gs, outs)
            return
        cur = torch.cuda.current_stream(self.device)
        self._comm.wait_stream(cur)
        with torch.cuda.stream(self._comm):
            self.reduce_scatter_chunk(c, gs, outs)
        for t in list(gs) + list(outs):
            t.record_stream(self._comm)

    def end_overlapped_reduce(self) -> List[torch.Tensor]:
        outs, _ = self._ov
        self._ov = None
        if self._comm is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm)
        return [o[: self.part.n_rows] for o in outs]

    # -- differentiable gather --------------------------------------------------------------
    def gather(self, *xs: torch.Tensor):
        """Differentiable halo all-gather: owned slices -> gathered tensors; the backward
        reduce-scatters the gradients back to the owned rows."""
        if not self.active:
            return xs if len(xs) > 1 else xs[0]
        out = _HaloGather.apply(self, *xs)
        return out if len(xs) > 1 else out[0]

    # round-1 names, kept for callers that time the two directions by hand
    def gather_pair(self, a, b, rec=None):
        return tuple(self.all_gather([a, b]))

    def reduce_pair(self, ga, gb, rec=None):
        return tuple(self.reduce_scatter([ga, gb]))


class _HaloGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, halo: HaloExchange, *xs):
        ctx.halo = halo
        ctx.rows = [x.shape[0] for x in xs]
        return tuple(halo.all_gather(xs))

    @staticmethod
    def backward(ctx, *grads):
        outs = ctx.halo.reduce_scatter([g.contiguous() for g in grads])
        # an operand that was passed padded gets a padded gradient back
        fixed = []
        for o, r in zip(outs, ctx.rows):
            if r != o.shape[0]:
                p = o.new_zeros((r,) + tuple(o.shape[1:]))
                p[: o.shape[0]].copy_(o)
                o = p
            fixed.append(o)
        return (None, *fixed)


# --------------------------------------------------------------------------------------- #
# distributed conv operators                                                               #
# --------------------------------------------------------------------------------------- #

def _col_chunks(halo: HaloExchange):
    """[(first column, number of columns)] of the column-side launches."""
    p = halo.part
    if not halo.active or p.chunks == 1:
        return [(0, p.n_cols)]
    w = halo.world * p.q
    return [(c * w, w) for c in range(p.chunks)]


def _fit_grad(g: torch.Tensor, rows: int) -> torch.Tensor:
    """Gradient of an operand that was given with `rows` rows (owned or padded)."""
    if g.shape[0] == rows:
        return g
    out = g.new_zeros((rows,) + tuple(g.shape[1:]))
    out[: g.shape[0]].copy_(g)
    return out


def dist_gt_forward(halo, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem_consume,
                    Q, K_own, V_own):
    """Halo all-gather of K, V + fused conv forward on the shard.  -> out, saved (for
    dist_gt_backward).  Plain function: the autograd Function below and callers that drive the
    two directions by hand (bench.py's device-timed step) share it."""
    from .operators import _native as N
    K, V = halo.all_gather([K_own, V_own])
    out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx,
                                   smem_consume, Q, K, V)
    return out, (row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, Q, K, V, attn)


def dist_gt_backward(halo, saved, smem_consume, grad_out, own_rows=None):
    """Row-side kernel, then the column-side kernel chunk by chunk with the reduce-scatter of
    dK, dV of chunk c running behind the kernel of chunk c+1.  -> dQ, dK_own, dV_own."""
    from .operators import _native as N
    row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, Q, K, V, attn = saved
    args = (row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem_consume, Q, K, V, attn,
            grad_out)
    bufs = N.gt_backward(*args, _phases=1)
    gq, gk, gv, ge = bufs
    if not halo.active:
        N.gt_backward(*args, _phases=2, _buffers=bufs)
        return gq, gk, gv
    if halo.backend == "p2p":  # the column side writes straight into peer-visible memory
        gk, gv = halo.p2p_grad_buffers([gk, gv])
        N.gt_backward(*args, _phases=2, _buffers=(gq, gk, gv, ge))
        gk_own, gv_own = halo.reduce_scatter([gk, gv])
        if own_rows is not None:
            gk_own, gv_own = _fit_grad(gk_own, own_rows[0]), _fit_grad(gv_own, own_rows[1])
        return gq, gk_own, gv_own
    halo.begin_overlapped_reduce([gk, gv])
    for c, (c0, nc) in enumerate(_col_chunks(halo)):
        N.gt_backward(*args, _phases=2, _buffers=bufs, _cols=(c0, nc, halo.chunk_nnz(col_ptr, c)))
        halo.reduce_chunk_async(c)
    gk_own, gv_own = halo.end_overlapped_reduce()
    if own_rows is not None:
        gk_own, gv_own = _fit_grad(gk_own, own_rows[0]), _fit_grad(gv_own, own_rows[1])
    return gq, gk_own, gv_own


def dist_gat_forward(halo, attn_row, attn_col_own, row_ptr, col_ind, col_ptr, row_ind, permute,
                     negative_slope, feat_own, attn_drop):
    from .operators import _native as N
    feat, ac = halo.all_gather([feat_own, attn_col_own])
    out, emax, esum, emask = N.gat_forward(attn_row, ac, row_ptr, col_ind, negative_slope, feat,
                                           attn_drop)
    return out, (row_ptr, col_ind, col_ptr, row_ind, permute, emax, esum, emask, feat, attn_row, ac)


def dist_gat_backward(halo, saved, negative_slope, attn_drop, grad_out, own_rows=None):
    """-> grad_attn_row, grad_attn_col_own, grad_feat_own."""
    from .operators import _native as N
    row_ptr, col_ind, col_ptr, row_ind, permute, emax, esum, emask, feat, attn_row, ac = saved
    args = (negative_slope, attn_drop, row_ptr, col_ind, col_ptr, row_ind, permute, emax, esum,
            emask, feat, attn_row, ac, grad_out)
    bufs = N.gat_backward(*args, _phases=1)
    gf, gr, gc, ge = bufs
    if not halo.active:
        N.gat_backward(*args, _phases=2, _buffers=bufs)
        return gr, gc, gf
    if halo.backend == "p2p":
        gf, gc = halo.p2p_grad_buffers([gf, gc])
        N.gat_backward(*args, _phases=2, _buffers=(gf, gr, gc, ge))
        gf_own, gc_own = halo.reduce_scatter([gf, gc])
        if own_rows is not None:
            gf_own, gc_own = _fit_grad(gf_own, own_rows[0]), _fit_grad(gc_own, own_rows[1])
        return gr, gc_own, gf_own
    halo.begin_overlapped_reduce([gf, gc])
    for c, (c0, nc) in enumerate(_col_chunks(halo)):
        N.gat_backward(*args, _phases=2, _buffers=bufs, _cols=(c0, nc, halo.chunk_nnz(col_ptr, c)))
        halo.reduce_chunk_async(c)
    gf_own, gc_own = halo.end_overlapped_reduce()
    if own_rows is not None:
        gf_own, gc_own = _fit_grad(gf_own, own_rows[0]), _fit_grad(gc_own, own_rows[1])
    return gr, gc_own, gf_own


class DistGTFunction(torch.autograd.Function):
    """Row-partitioned FusedGTFunction_hyper (operators/fused_gtconv.py:79-158 on a shard)."""

    @staticmethod
    def forward(ctx, halo, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem_consume,
                Q, K_own, V_own):
        out, saved = dist_gt_forward(halo, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx,
                                     smem_consume, Q, K_own, V_own)
        ctx.halo, ctx.smem = halo, smem_consume
        ctx.own_rows = (K_own.shape[0], V_own.shape[0])
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        gq, gk, gv = dist_gt_backward(ctx.halo, ctx.saved_tensors, ctx.smem, grad_out.contiguous(),
                                      ctx.own_rows)
        return (None,) * 9 + (gq, gk, gv)


class DistGATFunction(torch.autograd.Function):
    """Row-partitioned FusedGATFunction (operators/fused_gatconv.py:95-176 on a shard)."""

    @staticmethod
    def forward(ctx, halo, attn_row, attn_col_own, row_ptr, col_ind, col_ptr, row_ind, permute,
                negative_slope, feat_own, attn_drop):
        out, saved = dist_gat_forward(halo, attn_row, attn_col_own, row_ptr, col_ind, col_ptr,
                                      row_ind, permute, negative_slope, feat_own, attn_drop)
        ctx.halo, ctx.slope, ctx.drop = halo, negative_slope, attn_drop
        ctx.own_rows = (feat_own.shape[0], attn_col_own.shape[0])
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        gr, gc, gf = dist_gat_backward(ctx.halo, ctx.saved_tensors, ctx.slope, ctx.drop,
                                       grad_out.contiguous(), ctx.own_rows)
        return (None, gr, gc) + (None,) * 6 + (gf, None)


def GTConvFuse_hyper_dist(halo: HaloExchange, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx,
                          smem_consume, Q, K_own, V_own):
    """GTConvFuse_hyper (operators/fused_gtconv.py:51-76) on a row-partitioned shard: Q holds the
    rank's rows, K_own / V_own the rank's slice of the column-side operands ([n_rows, h, f] or
    already padded to [max_rows, h, f]); the index arrays describe the shard in the padded column
    space (``Partition.local_graph``).  Returns out [n_rows, h, f]; differentiable in Q, K_own, V_own."""
    return DistGTFunction.apply(halo, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx,
                                smem_consume, Q, K_own, V_own)


def GATConvFuse_dist(halo: HaloExchange, attn_row, attn_col_own, row_ptr, col_ind, col_ptr, row_ind,
                     permute, negative_slope, feat_own, attn_drop):
    """GATConvFuse (operators/fused_gatconv.py:5-28) on a row-partitioned shard."""
    return DistGATFunction.apply(halo, attn_row, attn_col_own, row_ptr, col_ind, col_ptr, row_ind,
                                 permute, negative_slope, feat_own, attn_drop)
