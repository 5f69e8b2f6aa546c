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


def test_balanced_work_lists_of_the_dense_kernels():
    """formats.balanced_lists: every graph (or (graph, key tile) item) exactly once, loads within a few
    per cent of each other on the PATTERN-shaped batch (round robin: 1.65x the mean on the busiest CTA)."""
    import numpy as np
    from dfgnn_b200.formats import balanced_lists
    bnn = graphs.pattern_like(batch=1024).batch_num_nodes().numpy()
    g, ptr, idx = balanced_lists(bnn, 148)
    assert g == 148 and ptr[0] == 0 and ptr[-1] == len(bnn) and sorted(idx.tolist()) == list(range(len(bnn)))
    tiles = np.where(bnn <= 128, 1.0, 3.0)  # a wide graph is two row tiles over twice the keys
    load = np.array([tiles[idx[ptr[c]:ptr[c + 1]]].sum() for c in range(g)])
    assert load.max() <= 1.15 * load.mean()
    g2, ptr2, idx2 = balanced_lists(bnn, 148, column_items=True)
    want = sorted([2 * b for b in range(len(bnn))] + [2 * b + 1 for b in range(len(bnn)) if bnn[b] > 128])
    assert sorted(idx2.tolist()) == want and ptr2[-1] == len(want)
    g3, ptr3, idx3 = balanced_lists(bnn[:5], 148)
    assert g3 == 5 and ptr3.tolist() == [0, 1, 2, 3, 4, 5]


def test_graph_store_round_trip(tmp_path):
    """graphs.save_graph / load_graph / GraphStore: identical edge lists, batch sizes and fingerprints;
    a damaged file is refused."""
    import numpy as np
    store = graphs.GraphStore(str(tmp_path))
    g = store.get("pattern_like", batch=6)
    g2 = store.get("pattern_like", batch=6)  # second call reads the file
    assert g2.sha256() == g.sha256() and g2.num_nodes() == g.num_nodes() and g2.batch_size == 6
    assert torch.equal(g2.batch_num_nodes(), g.batch_num_nodes())
    for a, b in zip(g.edges(), g2.edges()):
        assert torch.equal(a, b)
    c = graphs.constant_degree_graph(100, 5, seed=3)
    deg = torch.bincount(c.edges()[0], minlength=100)
    assert c.num_nodes() == 100 and int(deg.max()) <= 5 and int(deg.min()) >= 1
    path = str(tmp_path / "c.npz")
    graphs.save_graph(c, path)
    z = dict(np.load(path))
    z["dst"] = z["dst"].copy()
    z["dst"][0] += 1
    np.savez(path, **z)
    import pytest
    with pytest.raises(RuntimeError, match="sha256"):
        graphs.load_graph(path)
