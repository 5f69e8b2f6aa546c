"""Synthetic workload generators: deterministic, sorted by (row, col), unique."""
import torch

from dfgnn_b200 import graphs


def _check_sorted_unique(g):
    src, dst = g.edges()
    key = src * g.num_nodes() + dst
    assert bool((key[1:] > key[:-1]).all())
    assert int(src.min()) >= 0 and int(dst.max()) < g.num_nodes()


def test_generators_are_deterministic_and_canonical():
    for fn, kw in ((graphs.cora_like, {}), (graphs.arxiv_like, dict(scale=0.05)),
                   (graphs.pattern_like, dict(batch=8)), (graphs.pascalvoc_like, dict(batch=8)),
                   (graphs.reddit_like, dict(scale=0.02))):
        a, b = fn(**kw), fn(**kw)
        assert a.sha256() == b.sha256()
        _check_sorted_unique(a)


def test_shapes_follow_baseline_configs():
    g = graphs.cora_like()
    assert g.num_nodes() == 2708 and abs(g.num_edges() - 10556) < 200
    g = graphs.pattern_like(batch=32)
    bnn = g.batch_num_nodes()
    assert g.batch_size == 32 and int(bnn.sum()) == g.num_nodes()
    assert 50 <= int(bnn.min()) and int(bnn.max()) <= 186
    # block diagonal: every edge stays inside its graph
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(bnn, 0)])
    src, dst = g.edges()
    gid = torch.bucketize(src, offs[1:], right=True)
    assert bool(((dst >= offs[gid]) & (dst < offs[gid + 1])).all())
    deg = torch.bincount(src, minlength=g.num_nodes()).float()
    assert 45 < float(deg.mean()) < 57
    g = graphs.pascalvoc_like(batch=16)
    deg = torch.bincount(g.edges()[0], minlength=g.num_nodes()).float()
    assert 4.5 < float(deg.mean()) < 6.5


def test_conv_inputs_seeded():
    a, b = graphs.conv_inputs(10, 8, 5), graphs.conv_inputs(10, 8, 5)
    assert torch.equal(a.Q, b.Q) and torch.equal(a.dO, b.dO)
    assert a.Q.shape == (10, 1, 8) and a.attn_row.shape == (10, 1)
