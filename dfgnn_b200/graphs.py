"""Deterministic synthetic graphs shaped like the reference's benchmark datasets.

The reference loads cora / ogbn-arxiv / PATTERN / reddit / PascalVOC-SP through
DGL and OGB (``DFGNN/utils/util.py:41-148``, commented out in the snapshot);
neither the libraries nor the datasets exist offline, so every workload here is
a seeded synthetic graph with the node/edge counts of ``BASELINE.json:configs``
and the degree statistics of ``figure/graph_statistics/*.png``.

All generators emit a COO edge list **sorted by (row, col) with unique edges**
so that format construction (stable sort by row) is independent of tie-breaking
(SURVEY.md 8c).  ``row`` is the source node and the softmax runs over each row
(``DFGNN/layers/util.py:52-57``).

The ``Graph`` class is a duck-typed stand-in for the handful of ``dgl.DGLGraph``
methods the reference's preprocessing touches: ``edges()``, ``num_nodes()``,
``num_edges()``, ``batch_num_nodes()``, ``batch_size``, ``to(device)``.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch


class Graph:
    """Minimal COO graph container (the subset of DGLGraph used by preprocessing)."""

    def __init__(self, src: torch.Tensor, dst: torch.Tensor, num_nodes: int,
                 batch_num_nodes: Optional[torch.Tensor] = None, name: str = "graph",
                 num_cols: Optional[int] = None):
        assert src.shape == dst.shape and src.dim() == 1
        self._src = src
        self._dst = dst
        self._n = int(num_nodes)
        self._bnn = batch_num_nodes
        self.name = name
        # a row-partitioned shard is rectangular: num_nodes() local rows x num_cols columns
        self.num_cols = self._n if num_cols is None else int(num_cols)

    def edges(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._src, self._dst

    def num_nodes(self) -> int:
        return self._n

    def num_edges(self) -> int:
        return int(self._src.numel())

    def batch_num_nodes(self) -> torch.Tensor:
        if self._bnn is None:
            return torch.tensor([self._n], dtype=torch.int64)
        return self._bnn

    @property
    def batch_size(self) -> int:
        return 1 if self._bnn is None else int(self._bnn.numel())

    @property
    def device(self) -> torch.device:
        return self._src.device

    def to(self, device) -> "Graph":
        bnn = None if self._bnn is None else self._bnn.to(device)
        return Graph(self._src.to(device), self._dst.to(device), self._n, bnn, self.name,
                     self.num_cols)

    def sha256(self) -> str:
        """Fingerprint of the edge list (recorded in BASELINE.md once frozen)."""
        h = hashlib.sha256()
        h.update(self._src.cpu().to(torch.int64).numpy().tobytes())
        h.update(self._dst.cpu().to(torch.int64).numpy().tobytes())
        return h.hexdigest()

    def __repr__(self) -> str:
        return f"Graph({self.name}, N={self._n}, E={self.num_edges()}, batch={self.batch_size})"


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(int(seed))
    return g


def _lognormal_degrees(n: int, mean: float, std: float, dmin: int, dmax: int,
                       total: Optional[int], gen: torch.Generator) -> torch.Tensor:
    """Integer degrees ~ log-normal(mean, std) clamped to [dmin, dmax]; if
    ``total`` is given the degrees are rescaled so that they sum to it."""
    sigma2 = math.log(1.0 + (std / mean) ** 2)
    mu = math.log(mean) - 0.5 * sigma2
    z = torch.randn(n, generator=gen, dtype=torch.float64)
    deg = torch.exp(mu + math.sqrt(sigma2) * z)
    deg = deg.clamp(min=float(dmin), max=float(dmax))
    if total is not None:
        # rescale, floor, then hand the remainder to the largest fractional parts
        for _ in range(4):
            deg = (deg * (float(total) / float(deg.sum()))).clamp(min=float(dmin), max=float(dmax))
        fl = deg.floor()
        rem = int(total - int(fl.sum()))
        if rem > 0:
            frac = deg - fl
            frac[fl >= dmax] = -1.0
            idx = torch.topk(frac, min(rem, n)).indices
            fl[idx] += 1
        deg = fl
    else:
        deg = deg.round()
    return deg.clamp(min=float(dmin), max=float(dmax)).to(torch.int64)


def _sorted_unique_coo(row: torch.Tensor, col: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    key = row * n + col
    key = torch.unique(key)  # sorted ascending => sorted by (row, col)
    return torch.div(key, n, rounding_mode="floor"), key % n


def _uniform_columns(deg: torch.Tensor, n_cols: int, gen: torch.Generator,
                     col_offset: Optional[torch.Tensor] = None,
                     col_range: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """For each row i draw deg[i] columns uniformly (with replacement; caller dedupes)."""
    n = deg.numel()
    row = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg)
    u = torch.rand(row.numel(), generator=gen, dtype=torch.float64)
    if col_range is None:
        col = (u * n_cols).to(torch.int64).clamp_(max=n_cols - 1)
    else:
        r = col_range[row].to(torch.float64)
        col = (u * r).to(torch.int64)
        col = torch.minimum(col, col_range[row] - 1) + col_offset[row]
    return row, col


def full_graph(n: int, e_target: int, mean: float, std: float, dmin: int, dmax: int,
               seed: int, name: str) -> Graph:
    """Full graph with log-normal row degrees and uniform random columns."""
    gen = _gen(seed)
    deg = _lognormal_degrees(n, mean, std, dmin, dmax, e_target, gen)
    row, col = _uniform_columns(deg, n, gen)
    row, col = _sorted_unique_coo(row, col, n)
    return Graph(row, col, n, None, name)


def batched_graph(batch: int, nodes_mean: float, nodes_std: float, nodes_min: int, nodes_max: int,
                  deg_mean: float, deg_std: float, deg_min: int, deg_max_cap: Optional[int],
                  seed: int, name: str) -> Graph:
    """Block-diagonal batch of ``batch`` small graphs (a DGL batched graph).

    Nodes per graph ~ N(nodes_mean, nodes_std) clamped; per-row degree
    ~ N(deg_mean, deg_std) clamped to [deg_min, min(n_g-1, cap)]; columns drawn
    uniformly without replacement inside the row's own graph.
    """
    gen = _gen(seed)
    ng = torch.randn(batch, generator=gen, dtype=torch.float64) * nodes_std + nodes_mean
    ng = ng.round().clamp(min=nodes_min, max=nodes_max).to(torch.int64)
    offs = torch.zeros(batch + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(ng, 0)
    n = int(offs[-1])
    gid = torch.repeat_interleave(torch.arange(batch, dtype=torch.int64), ng)
    n_of_row = ng[gid]
    deg = torch.randn(n, generator=gen, dtype=torch.float64) * deg_std + deg_mean
    hi = n_of_row - 1
    if deg_max_cap is not None:
        hi = hi.clamp(max=deg_max_cap)
    deg = torch.minimum(deg.round().clamp(min=deg_min).to(torch.int64), hi.clamp(min=deg_min))
    density = float(deg.sum()) / float((n_of_row).sum())
    if density > 0.10:
        # dense-ish blocks: exact sampling without replacement by ranking uniforms
        rows, cols = [], []
        for b in range(batch):
            nb = int(ng[b])
            lo = int(offs[b])
            u = torch.rand(nb, nb, generator=gen)
            rank = u.argsort(dim=1).argsort(dim=1)
            mask = rank < deg[lo:lo + nb].unsqueeze(1)
            r, c = mask.nonzero(as_tuple=True)  # row-major => sorted by (row, col)
            rows.append(r + lo)
            cols.append(c + lo)
        row = torch.cat(rows)
        col = torch.cat(cols)
    else:
        row, col = _uniform_columns(deg, n, gen, col_offset=offs[:-1][gid], col_range=n_of_row)
        row, col = _sorted_unique_coo(row, col, n)
    return Graph(row, col, n, ng, name)


# --------------------------------------------------------------------------- #
# The five BASELINE.json configs.  ``scale`` < 1 shrinks  #
# the node/graph count for tests; 1.0 is the benchmark size.                   #
# --------------------------------------------------------------------------- #

def cora_like(scale: float = 1.0, seed: int = 1001) -> Graph:
    n = max(8, int(round(2708 * scale)))
    e = max(8, int(round(10556 * scale)))
    return full_graph(n, e, 3.90, 5.23, 1, min(168, n - 1), seed, "cora-shaped")


def arxiv_like(scale: float = 1.0, seed: int = 1002) -> Graph:
    n = max(8, int(round(169343 * scale)))
    e = max(8, int(round(1166243 * scale)))
    return full_graph(n, e, 6.89, 8.88, 0, min(436, n - 1), seed, "arxiv-shaped")


def pattern_like(batch: int = 1024, seed: int = 1003) -> Graph:
    return batched_graph(batch, 118.91, 21.07, 50, 186, 51.13, 11.13, 1, None, seed,
                         "PATTERN-shaped")


def reddit_like(scale: float = 1.0, seed: int = 1004) -> Graph:
    n = max(64, int(round(232965 * scale)))
    e = int(round(114615892 * scale * scale)) if scale < 1.0 else 114615892
    mean = e / n
    return full_graph(n, e, mean, mean * 1.6, 1, min(21657, n - 1), seed, "reddit-shaped")


def pascalvoc_like(batch: int = 1024, seed: int = 1005) -> Graph:
    return batched_graph(batch, 479.25, 16.8, 150, 500, 5.65, 1.21, 1, 31, seed,
                         "PascalVOC-SP-shaped")


def constant_degree_graph(num_nodes: int, num_neigh: int, seed: int = 0) -> Graph:
    """Every row has ``num_neigh`` random neighbours (``DFGNN/utils/graph_generate.py:20-27``; duplicates
    removed so that the edge list is a set, like every other generator here)."""
    gen = _gen(seed)
    row = torch.arange(num_nodes, dtype=torch.int64).repeat_interleave(num_neigh)
    col = torch.randint(0, max(1, num_nodes - 1), (row.numel(),), generator=gen, dtype=torch.int64)
    row, col = _sorted_unique_coo(row, col, num_nodes)
    return Graph(row, col, num_nodes, None, f"constant-degree-{num_neigh}")


# --------------------------------------------------------------------------- #
# On-disk graph store: one .npz per graph -- int32 COO,     #
# node count, graph sizes of a batch, name and the sha256 of the edge list --    #
# so that a workload is generated once and every later run (and every rank)      #
# reads identical bytes.  The reference's dataset loaders are commented out      #
# (DFGNN/utils/util.py:41-148) and need DGL / OGB downloads; the store holds     #
# the synthetic graphs of the same shapes.                                       #
# --------------------------------------------------------------------------- #

def save_graph(g: Graph, path: str) -> str:
    """-> sha256 of the edge list (also stored in the file and checked by ``load_graph``)."""
    import numpy as np
    src, dst = g.edges()
    sha = g.sha256()
    bnn = g.batch_num_nodes().cpu().numpy().astype(np.int32) if g.batch_size > 1 else np.zeros(0, np.int32)
    dt = np.int32 if max(g.num_nodes(), g.num_cols) < 2 ** 31 else np.int64
    np.savez(path, src=src.cpu().numpy().astype(dt), dst=dst.cpu().numpy().astype(dt),
             num_nodes=np.int64(g.num_nodes()), num_cols=np.int64(g.num_cols), batch_num_nodes=bnn,
             name=np.array(g.name), sha256=np.array(sha))
    return sha


def load_graph(path: str) -> Graph:
    import numpy as np
    with np.load(path if path.endswith(".npz") else path + ".npz", allow_pickle=False) as z:
        bnn = torch.from_numpy(z["batch_num_nodes"].astype(np.int64)) if z["batch_num_nodes"].size else None
        g = Graph(torch.from_numpy(z["src"].astype(np.int64)), torch.from_numpy(z["dst"].astype(np.int64)),
                  int(z["num_nodes"]), bnn, str(z["name"]), int(z["num_cols"]))
        want = str(z["sha256"])
    if g.sha256() != want:
        raise RuntimeError(f"{path}: edge list does not match its recorded sha256 (corrupt or truncated file)")
    return g


class GraphStore:
    """Directory of generated graphs: ``store.get("pattern_like", batch=64)`` generates on the first
    call and loads the file afterwards."""

    def __init__(self, root: str):
        import os
        self.root = root
        os.makedirs(root, exist_ok=True)

    def path(self, fn: str, **kw) -> str:
        import os
        tag = "_".join(f"{k}-{v}" for k, v in sorted(kw.items()))
        return os.path.join(self.root, fn + ("_" + tag if tag else "") + ".npz")

    def get(self, fn: str, **kw) -> Graph:
        import os
        path = self.path(fn, **kw)
        if os.path.exists(path):
            return load_graph(path)
        g = globals()[fn](**kw)
        save_graph(g, path)
        return g


@dataclass
class ConvInputs:
    """Seeded operands for one conv call (features seed+1, dO seed+2)."""
    Q: torch.Tensor
    K: torch.Tensor
    V: torch.Tensor
    dO: torch.Tensor
    attn_row: torch.Tensor
    attn_col: torch.Tensor


def conv_inputs(n: int, dim: int, seed: int, heads: int = 1) -> ConvInputs:
    g1 = _gen(seed + 1)
    g2 = _gen(seed + 2)
    Q = torch.randn(n, heads, dim, generator=g1) * (dim ** -0.5)
    K = torch.randn(n, heads, dim, generator=g1)
    V = torch.randn(n, heads, dim, generator=g1)
    ar = torch.randn(n, heads, generator=g1)
    ac = torch.randn(n, heads, generator=g1)
    dO = torch.randn(n, heads, dim, generator=g2)
    return ConvInputs(Q, K, V, dO, ar, ac)
