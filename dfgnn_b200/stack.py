"""The callers either side of the conv: the multi-layer model loop and the per-batch format
construction of the reference's training scripts.

* ``GTStack`` is the layer loop of ``script/train/train_full_graph_timing.py:14-35`` (``Net``:
  input projection, L x ``SparseMHA_forward``, output projection + log-softmax) and, with
  ``pool=True``, of ``script/train/train_gtconv.py:51-77`` (``GTModel``: sum pooling per graph of
  the batch before the predictor).  The conv inside every layer is the fused operator.
* ``GraphedTrainStep`` captures ONE training step of such a stack -- forward, loss, backward and
  the optimizer update, i.e. 3 conv kernels x L layers plus the dense layers around them -- in a
  single CUDA graph.  The index formats (CSR / CSC / block plan) are built once and stay resident
  across layers and steps, like the reference's ``params = preprocess_Hyper_fw_bw(g, True)``
  (``train_full_graph_timing.py:57``); a step is then one graph launch instead of ~40 Python-level
  operator calls per layer.
* ``FormatPrefetcher`` builds the formats of batch t+1 on a side stream (a worker thread owns the
  host-side waits of the format kernels) while the conv of batch t runs -- the reference builds
  them inline inside the training loop (``train_batch_graph_timing.py:170``) and times that cost
  separately (l.115-143).
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Iterable, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layers.GT import SparseMHA_forward


class GTStack(nn.Module):
    def __init__(self, num_layers: int, in_dim: int, num_hidden: int, num_classes: int,
                 num_heads: int = 1, pool: bool = False, fused_projection: bool = False):
        super().__init__()
        self.num_layers = num_layers
        self.pool = pool
        self.input_proj = nn.Linear(in_dim, num_hidden)
        self.layers = nn.ModuleList(SparseMHA_forward(num_hidden, num_hidden, num_heads)
                                    for _ in range(num_layers))
        for layer in self.layers:  # q, k, v of every layer by the tcgen05 projection kernel
            layer.fused_projection = fused_projection
        self.output_proj = nn.Linear(num_hidden, num_classes)

    def forward(self, params, h, fuse: bool = True, graph_ids: Optional[torch.Tensor] = None,
                num_graphs: int = 0):
        h = self.input_proj(h)
        for layer in self.layers:
            h = layer(params, h, fuse)
        if self.pool:  # dglnn.SumPooling (train_gtconv.py:69, 75): one row per graph of the batch
            pooled = h.new_zeros((num_graphs, h.shape[1]))
            h = pooled.index_add_(0, graph_ids, h)
            return self.output_proj(h)
        return F.log_softmax(self.output_proj(h), dim=-1)


class GraphedTrainStep:
    """One training step (forward + loss + backward + optimizer update) of ``model`` on fixed
    ``params`` (resident index formats) as a CUDA graph.  ``step(x, y)`` copies the batch into the
    static input buffers, replays the graph and returns the (static) loss tensor."""

    def __init__(self, model: nn.Module, params, optimizer: torch.optim.Optimizer,
                 loss_fn: Callable, x: torch.Tensor, y: torch.Tensor, warmup: int = 3, **fwd_kwargs):
        self.model, self.params, self.opt, self.loss_fn = model, params, optimizer, loss_fn
        self.x = x.clone()
        self.y = y.clone()
        self.kw = fwd_kwargs
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up off the capture stream: allocator, autograd, plan caches
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._eager(zero=False)

    def _eager(self, zero: bool = True):
        if zero:
            self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(self.params, self.x, True, **self.kw), self.y)
        loss.backward()
        self.opt.step()
        return loss.detach()

    def step(self, x: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss


class FormatPrefetcher:
    """Index formats of the NEXT batch built on a side stream while the current batch computes.

        for g, params in FormatPrefetcher(preprocess_Hyper_fw_bw, device).iterate(batches):
            out = layer(params, feats_of(g), fuse=True)

    ``preprocess`` runs in a worker thread under its own CUDA stream (its small validation
    read-backs block only that thread); the consumer's stream waits on the event recorded behind
    the batch's format kernels, never on the host."""

    def __init__(self, preprocess: Callable, device, depth: int = 2):
        self.preprocess = preprocess
        self.device = torch.device(device)
        self.depth = max(1, depth)
        self.stream = torch.cuda.Stream(self.device)

    def _tensors(self, obj):
        if isinstance(obj, torch.Tensor):
            yield obj
        elif isinstance(obj, (tuple, list)):
            for o in obj:
                yield from self._tensors(o)
        elif hasattr(obj, "__dict__"):
            for o in vars(obj).values():
                if isinstance(o, torch.Tensor):
                    yield o

    def iterate(self, graphs: Iterable):
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        _END = object()

        def work():
            try:
                torch.cuda.set_device(self.device)
                for g in graphs:
                    with torch.cuda.stream(self.stream):
                        g = g.to(self.device) if hasattr(g, "to") else g
                        params = self.preprocess(g)
                        ev = torch.cuda.Event()
                        ev.record(self.stream)
                    q.put((g, params, ev))
                q.put(_END)
            except BaseException as exc:  # surface worker failures in the consumer
                q.put(exc)

        t = threading.Thread(target=work, daemon=True)
        t.start()
        while True:
            item = q.get()
            if item is _END:
                break
            if isinstance(item, BaseException):
                raise item
            g, params, ev = item
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for ten in self._tensors(params):
                if ten.is_cuda:
                    ten.record_stream(cur)
            yield g, params
        t.join()
