"""The callers either side of the conv: the multi-layer loop as one CUDA graph and
the side-stream format construction."""
import copy

import pytest
import torch
import torch.nn.functional as F

from dfgnn_b200 import graphs
from dfgnn_b200.layers import preprocess_Hyper_fw_bw
from dfgnn_b200.stack import FormatPrefetcher, GraphedTrainStep, GTStack
from dfgnn_b200.utils import check_correct

from .helpers import assert_close

pytestmark = pytest.mark.gpu


def test_stack_fused_matches_nonfused(cuda):
    torch.manual_seed(0)
    g = graphs.pattern_like(batch=4).to(cuda)
    params = preprocess_Hyper_fw_bw(g)
    model = GTStack(3, 32, 64, 7).to(cuda).eval()
    x = torch.randn(g.num_nodes(), 32, device=cuda)
    with torch.no_grad():
        a = model(params, x, fuse=True)
        b = model(params, x, fuse=False)
    assert a.shape == (g.num_nodes(), 7)
    assert check_correct(b, a)
    assert_close("stack out", a, b, rtol=1e-3, atol=1e-5)
    # pooled variant (GTModel, train_gtconv.py:51-77)
    pooled = GTStack(2, 32, 64, 1, pool=True).to(cuda).eval()
    gid = torch.repeat_interleave(torch.arange(g.batch_size, device=cuda), g.batch_num_nodes())
    with torch.no_grad():
        y = pooled(params, x, True, graph_ids=gid, num_graphs=g.batch_size)
        y2 = pooled(params, x, False, graph_ids=gid, num_graphs=g.batch_size)
    assert y.shape == (g.batch_size, 1)
    assert_close("pooled stack out", y, y2, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("layers", [2, 8])
def test_graphed_train_step_equals_eager_steps(cuda, layers):
    """One CUDA graph = forward + loss + backward + SGD update of the whole stack: after the same
    number of steps on the same batches the weights equal those of the eager loop."""
    torch.manual_seed(1)
    g = graphs.pattern_like(batch=6).to(cuda)
    params = preprocess_Hyper_fw_bw(g)
    n = g.num_nodes()
    eager = GTStack(layers, 32, 64, 5).to(cuda).train()
    graphed = copy.deepcopy(eager)
    xs = [torch.randn(n, 32, device=cuda) for _ in range(3)]
    ys = [torch.randint(0, 5, (n,), device=cuda) for _ in range(3)]
    opt_e = torch.optim.SGD(eager.parameters(), lr=0.05)
    opt_g = torch.optim.SGD(graphed.parameters(), lr=0.05)
    # GraphedTrainStep warms up with real updates: give the eager model the same ones
    step = GraphedTrainStep(graphed, params, opt_g, F.nll_loss, xs[0], ys[0], warmup=2)
    for _ in range(2):
        opt_e.zero_grad(set_to_none=True)
        F.nll_loss(eager(params, xs[0], True), ys[0]).backward()
        opt_e.step()
    losses_e, losses_g = [], []
    for x, y in zip(xs, ys):
        opt_e.zero_grad(set_to_none=True)
        loss = F.nll_loss(eager(params, x, True), y)
        loss.backward()
        opt_e.step()
        losses_e.append(float(loss))
        losses_g.append(float(step.step(x, y)))
    assert losses_g == pytest.approx(losses_e, rel=1e-5, abs=1e-6)
    for (name, pe), pg in zip(eager.named_parameters(), graphed.parameters()):
        assert_close(name, pg, pe, rtol=1e-4, atol=1e-6)


def test_format_prefetcher_overlaps_and_matches_inline(cuda):
    batches = [graphs.pattern_like(batch=3, seed=40 + i) for i in range(5)]
    seen = 0
    for i, (g, params) in enumerate(FormatPrefetcher(preprocess_Hyper_fw_bw, cuda).iterate(batches)):
        inline = preprocess_Hyper_fw_bw(batches[i].to(cuda))
        assert g.num_nodes() == batches[i].num_nodes()
        for name, a, b in zip(("rows", "row_ptr", "col_ind", "val", "col_ptr", "row_ind", "val_idx"),
                              params[1:8], inline[1:8]):
            assert torch.equal(a, b), name
        assert params[8] == inline[8]
        assert getattr(params[2], "_dfgnn_blocks", None) is not None   # the block plan travels too
        seen += 1
    assert seen == len(batches)
    # a failing preprocess surfaces in the consumer
    def boom(g):
        raise ValueError("bad batch")
    with pytest.raises(ValueError, match="bad batch"):
        list(FormatPrefetcher(boom, cuda).iterate(batches[:1]))
