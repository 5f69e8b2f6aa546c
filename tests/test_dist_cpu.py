"""Host-side partition / halo-exchange logic at world_size 2 on the gloo backend.

Each rank runs the conv maths of ITS shard with the CPU oracle (this is a test:
the oracle is the checker for the index bookkeeping, not a product path) on
operands exchanged through dfgnn_b200.dist.HaloExchange, and rank 0 compares the
stitched result with the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dfgnn_b200 import dist as ddist
from dfgnn_b200 import graphs
from oracle import cpu_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = graphs.arxiv_like(0.01) if kind.startswith("row") else graphs.pattern_like(batch=5)
        n, dim = g.num_nodes(), 16
        X = graphs.conv_inputs(n, dim, 5)
        part = ddist.make_partition(g, world, rank, chunks=3 if kind == "row-chunked" else 1)
        assert part.kind == ("row" if kind.startswith("row") else "by-graph")
        halo = ddist.HaloExchange(part, "cpu", world)
        lg = part.local_graph
        src, dst = lg.edges()
        # the oracle builder is square: build on max(rows, cols) and cut the row pointer
        rp, ci, _, _ = O.coo_to_csr(src, dst, max(lg.num_nodes(), lg.num_cols))
        rp = rp[: lg.num_nodes() + 1]
        K, V = halo.gather_pair(X.K[part.col_owned].contiguous(), X.V[part.col_owned].contiguous())
        assert K.shape[0] == part.n_cols
        Q = X.Q[part.row_slice]
        out, attn = _rect_forward(rp, ci, Q, K, V)
        # backward: partial column-side grads over all padded columns, then reduce
        pad = np.full(max(0, part.n_cols - lg.num_nodes()), rp[-1], np.int32)
        cp, ri, vi = O.csr_to_csc(np.concatenate([rp, pad]), ci, part.n_cols)
        dO = X.dO[part.row_slice]
        dQ, dK, dV = _rect_backward(rp, ci, cp, ri, vi, Q, K, V, attn, dO, part.n_cols)
        gK, gV = halo.reduce_pair(torch.from_numpy(dK), torch.from_numpy(dV))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), out=out, dQ=dQ, dK=gK.numpy(), dV=gV.numpy(),
                 lo=part.row_slice.start, hi=part.row_slice.stop)
    finally:
        dist.destroy_process_group()


def _rect_forward(rp, ci, Q, K, V):
    """Oracle forward on an m x n shard: the C oracle indexes K/V by column id only."""
    Qn = np.ascontiguousarray(Q.numpy(), np.float64)
    m, h, f = Qn.shape
    Kn, Vn = K.numpy().astype(np.float64), V.numpy().astype(np.float64)
    out = np.zeros((m, h, f))
    attn = np.zeros((h, len(ci)))
    for i in range(m):
        lb, hb = rp[i], rp[i + 1]
        if hb == lb:
            continue
        s = np.einsum("hd,ehd->eh", Qn[i], Kn[ci[lb:hb]])
        p = np.exp(s - s.max(0))
        p /= p.sum(0)
        attn[:, lb:hb] = p.T
        out[i] = np.einsum("eh,ehd->hd", p, Vn[ci[lb:hb]])
    return out, attn


def _rect_backward(rp, ci, cp, ri, vi, Q, K, V, attn, dO, n_cols):
    Qn, Kn, Vn, g = (t.numpy().astype(np.float64) for t in (Q, K, V, dO))
    m, h, f = Qn.shape
    rows = np.repeat(np.arange(m), np.diff(rp))
    p = attn.T  # [E, h]
    dA = np.einsum("ehd,ehd->eh", g[rows], Vn[ci])
    t = dA * p
    s = np.zeros((m, h))
    np.add.at(s, rows, t)
    dS = t - s[rows] * p
    dQ = np.zeros_like(Qn)
    np.add.at(dQ, rows, dS[:, :, None] * Kn[ci])
    dK = np.zeros((n_cols, h, f))
    dV = np.zeros((n_cols, h, f))
    np.add.at(dK, ci, dS[:, :, None] * Qn[rows])
    np.add.at(dV, ci, p[:, :, None] * g[rows])
    return dQ, dK, dV


@pytest.mark.parametrize("kind", ["row", "row-chunked", "by-graph"])
def test_two_rank_partition_matches_single_process(tmp_path, kind):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), kind, str(tmp_path)), nprocs=world, join=True)
    g = graphs.arxiv_like(0.01) if kind.startswith("row") else graphs.pattern_like(batch=5)
    n = g.num_nodes()
    X = graphs.conv_inputs(n, 16, 5)
    src, dst = g.edges()
    rp, ci, _, _ = O.coo_to_csr(src, dst, n)
    cp, ri, vi = O.csr_to_csc(rp, ci, n)
    out, attn = O.gt_forward(rp, ci, None, X.Q, X.K, X.V, dtype=np.float64)
    dQ, dK, dV, _ = O.gt_backward(rp, ci, cp, ri, vi, X.Q, X.K, X.V, attn, X.dO, dtype=np.float64)
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"r{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        covered += hi - lo
        for name, got, want in (("out", z["out"], out), ("dQ", z["dQ"], dQ), ("dK", z["dK"], dK),
                                ("dV", z["dV"], dV)):
            assert np.allclose(got, want[lo:hi], rtol=1e-9, atol=1e-11), (kind, r, name)
    assert covered == n


def test_row_bounds_balance_edges():
    g = graphs.reddit_like(0.03)
    deg = torch.bincount(g.edges()[0], minlength=g.num_nodes())
    for world in (2, 4, 8):
        b = ddist.row_bounds(deg, world)
        assert int(b[0]) == 0 and int(b[-1]) == g.num_nodes() and bool((b[1:] >= b[:-1]).all())
        per = torch.stack([deg[int(b[r]):int(b[r + 1])].sum() for r in range(world)]).float()
        assert float(per.max() / per.mean()) < 1.15
    part = ddist.make_partition(g, 4, 1)
    src, dst = part.local_graph.edges()
    assert int(dst.max()) < part.n_cols == 4 * part.max_rows
    assert part.local_graph.num_nodes() == part.n_rows == part.row_slice.stop - part.row_slice.start


def _autograd_worker(rank, world, port, chunks, out_dir):
    """HaloExchange.gather is differentiable: its backward is the reduce-scatter."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = graphs.arxiv_like(0.01)
        n = g.num_nodes()
        part = ddist.make_partition(g, world, rank, chunks=chunks)
        halo = ddist.HaloExchange(part, "cpu", world)
        gen = torch.Generator().manual_seed(11)
        X = torch.randn(n, 2, 4, generator=gen, dtype=torch.float64)
        a = torch.randn(n, 2, generator=gen, dtype=torch.float64)
        W = torch.randn(world, part.n_cols, 2, 4, generator=gen, dtype=torch.float64)  # per-rank loss weights
        x_own = X[part.col_owned].clone().requires_grad_()
        a_pad = halo.pad(a[part.col_owned].clone()).requires_grad_()  # an operand given already padded
        Xg, ag = halo.gather(x_own, a_pad)
        idx = part.padded_index(torch.arange(n))
        assert torch.equal(Xg[idx], X) and torch.equal(ag[idx], a)
        ((Xg * W[rank]).sum() + (ag * W[rank][:, :, 0]).sum()).backward()
        want_x = W.sum(0)[idx][part.col_owned]
        want_a = W.sum(0)[:, :, 0][idx][part.col_owned]
        assert x_own.grad.shape == x_own.shape and a_pad.grad.shape == a_pad.shape
        assert torch.allclose(x_own.grad, want_x, rtol=1e-12, atol=1e-12)
        assert torch.allclose(a_pad.grad[: part.n_rows], want_a, rtol=1e-12, atol=1e-12)
        assert float(a_pad.grad[part.n_rows:].abs().sum()) == 0.0
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chunks", [1, 2])
def test_halo_gather_backward_is_the_reduce_scatter(tmp_path, chunks):
    world = 2
    mp.spawn(_autograd_worker, args=(world, _free_port(), chunks, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_padded_index_is_a_bijection_onto_the_owned_slots():
    g = graphs.arxiv_like(0.02)
    n = g.num_nodes()
    for world, chunks in ((2, 1), (4, 2), (8, 4)):
        parts = [ddist.make_partition(g, world, r, chunks=chunks) for r in range(world)]
        p0 = parts[0]
        idx = p0.padded_index(torch.arange(n))
        assert idx.unique().numel() == n and int(idx.max()) < p0.n_cols == world * p0.max_rows
        assert p0.max_rows % chunks == 0
        q = p0.q
        for r, p in enumerate(parts):
            own = idx[p.col_owned]
            loc = torch.arange(p.n_rows)
            assert torch.equal(own, (loc // q) * (world * q) + r * q + loc % q)
            assert torch.equal(p.padded_index(torch.arange(n)), idx)
