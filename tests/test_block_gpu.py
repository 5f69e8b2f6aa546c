"""Graph-resident GT kernels (csrc/block_gt.cuh) for block-diagonal batches: one CTA per graph,
K/V (dO/Q) of the graph staged in shared memory by TMA bulk copies.  Checked against the fp64 CPU
oracle and against the general row-block kernels on the same inputs; tolerance 1e-4 relative /
1e-5 absolute (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from dfgnn_b200 import _lib, formats, graphs
from dfgnn_b200.layers import preprocess_Hyper, preprocess_Hyper_fw_bw
from dfgnn_b200.operators import GTConvFuse_hyper
from dfgnn_b200.operators import _native as N
from oracle import cpu_oracle as O

from .helpers import assert_close, assert_close_bulk

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_block_kernels():
    """Mode 2: the block kernels run whenever the stage fits (the automatic choice is tuned for
    speed, these tests are about correctness)."""
    prev = _lib.lib().dfgnn_set_block_mode(2)
    yield
    _lib.lib().dfgnn_set_block_mode(prev)


def _batch(kind):
    if kind == "pattern":      # largest graph <= 159 nodes: 16 warps per CTA in all three kernels
        return graphs.batched_graph(8, 110.0, 20.0, 50, 150, 51.0, 11.0, 1, None, 3, "pattern-small")
    if kind == "pattern-max":  # graphs of up to 186 nodes: the backward kernels run with 8 warps
        return graphs.batched_graph(6, 170.0, 20.0, 120, 186, 51.0, 11.0, 1, None, 4, "pattern-max")
    if kind == "ragged":       # 1-node graphs, rows without edges, sparse and dense graphs mixed
        g = graphs.batched_graph(40, 30.0, 40.0, 1, 186, 9.0, 8.0, 0, None, 5, "ragged")
        return g
    raise KeyError(kind)


def _oracle(g, X):
    src, dst = g.edges()
    n = g.num_nodes()
    rp, ci, rows, perm = O.coo_to_csr(src, dst, n)
    cp, ri, vi = O.csr_to_csc(rp, ci, n)
    out, attn = O.gt_forward(rp, ci, None, X.Q, X.K, X.V, dtype=np.float64)
    dQ, dK, dV, _ = O.gt_backward(rp, ci, cp, ri, vi, X.Q, X.K, X.V, attn, X.dO, dtype=np.float64)
    return out, attn, dQ, dK, dV


@pytest.mark.parametrize("kind,dim", [("pattern", 128), ("pattern", 64), ("pattern", 32), ("pattern-max", 128),
                                      ("ragged", 128), ("ragged", 64)])
def test_block_kernels_match_oracle_and_general_kernels(cuda, kind, dim):
    g = _batch(kind)
    n = g.num_nodes()
    X = graphs.conv_inputs(n, dim, 17)
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(cuda))
    plan = getattr(row_ptr, "_dfgnn_blocks", None)
    assert plan is not None and plan.n_blocks == g.batch_size
    assert plan.max_nodes == int(g.batch_num_nodes().max())
    assert plan.algorithm(n, col_ind.numel(), 1, dim, True) == 1, "the batch should run on the block kernels"
    Q, K, V, dO = (t.to(cuda) for t in (X.Q, X.K, X.V, X.dO))
    out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    assert _lib.last_kernel(0) == "gt_block_fwd_kernel"
    gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO)
    assert _lib.last_kernel(1) == "gt_block_bwd_row_kernel" and _lib.last_kernel(2) == "gt_block_bwd_col_kernel"
    inf = N.gt_hyper_inference(row_ptr, col_ind, rows, val, smem, Q, K, V)[0]
    assert torch.equal(inf, out)
    o64, a64, dQ, dK, dV = _oracle(g, X)
    assert_close("out", out, o64)
    assert_close("attn_edge", attn, a64)
    assert_close("grad_Q", gq, dQ)
    assert_close("grad_K", gk, dK)
    assert_close("grad_V", gv, dV)
    # the same CSR without the plan runs the general kernels: same numbers within tolerance
    rp2 = row_ptr.clone()
    out2, attn2 = N.gt_hyper_forward(rp2, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    assert _lib.last_kernel(0) == "dot_fwd_kernel"
    gq2, gk2, gv2 = N.gt_backward(rp2, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn2, dO)
    for name, a, b in (("out", out, out2), ("attn", attn, attn2), ("gq", gq, gq2), ("gk", gk, gk2), ("gv", gv, gv2)):
        assert_close("block vs general " + name, a, b)


def test_block_path_under_autograd_and_weighted_scores(cuda):
    g = _batch("pattern")
    n = g.num_nodes()
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(cuda))
    X = graphs.conv_inputs(n, 128, 19)
    Q, K, V = (t.to(cuda).requires_grad_() for t in (X.Q, X.K, X.V))
    out = GTConvFuse_hyper(rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    out.backward(X.dO.to(cuda))
    # the plan travels with row_ptr through save_for_backward
    assert _lib.last_kernel(1) == "gt_block_bwd_row_kernel"
    o64, a64, dQ, dK, dV = _oracle(g, X)
    assert_close("out", out, o64)
    assert_close("dQ", Q.grad, dQ)
    assert_close("dK", K.grad, dK)
    assert_close("dV", V.grad, dV)
    # weighted scores through the block kernels vs fp64 autograd
    gen = torch.Generator().manual_seed(3)
    w = (0.5 + torch.rand(col_ind.numel(), generator=gen)).to(cuda)
    Qd, Kd, Vd = (t.detach() for t in (Q, K, V))
    o, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, w, col_ptr, row_ind, val_idx, smem, Qd, Kd, Vd)
    assert _lib.last_kernel(0) == "gt_block_fwd_kernel"
    gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, w, col_ptr, row_ind, val_idx, smem, Qd, Kd, Vd, attn,
                               X.dO.to(cuda))
    r, c = A.row.long(), A.col.long()   # sorted by (row, col): CSR order
    Q6, K6, V6 = (t[:, 0].double().requires_grad_() for t in (Qd, Kd, Vd))
    s = (Q6[r] * K6[c]).sum(-1) * w.double()
    mx = torch.full((n,), -1e300, dtype=torch.float64, device=cuda).scatter_reduce(0, r, s, "amax")
    ex = torch.exp(s - mx[r])
    p = ex / torch.zeros(n, dtype=torch.float64, device=cuda).index_add(0, r, ex)[r]
    ref = torch.zeros_like(V6).index_add(0, r, p[:, None] * V6[c])
    ref.backward(X.dO.to(cuda)[:, 0].double())
    assert_close("weighted out", o[:, 0], ref.detach())
    assert_close("weighted dQ", gq[:, 0], Q6.grad)
    assert_close("weighted dK", gk[:, 0], K6.grad)
    assert_close("weighted dV", gv[:, 0], V6.grad)


def test_block_plan_validation_and_fallbacks(cuda):
    g = _batch("pattern")
    gd = g.to(cuda)
    row_ptr, col_ind, rows, val, smem = preprocess_Hyper(gd)
    assert getattr(row_ptr, "_dfgnn_blocks", None) is not None
    # wrong node ranges: an edge leaves its block
    bnn = g.batch_num_nodes().clone()
    bnn[0] -= 5
    bnn[1] += 5
    with pytest.raises(RuntimeError, match="not block diagonal"):
        formats.block_plan(bnn, row_ptr, col_ind)
    with pytest.raises(RuntimeError, match="sums to"):
        formats.block_plan(bnn[:-1], row_ptr, col_ind)
    # sizes the stage cannot hold (VOC-shaped graphs: ~480 nodes x 128 floats x 2 > 227 KB), several
    # heads, or short rows fall back to the general kernels
    plan = row_ptr._dfgnn_blocks
    n, nnz = g.num_nodes(), col_ind.numel()
    assert plan.supported(n, nnz, 1, 128) and not plan.supported(n, nnz, 2, 64)
    assert not plan.supported(n, nnz, 1, 48)
    big = formats.BlockPlan(plan.blk_ptr, plan.n_blocks, 480)
    assert not big.supported(n, nnz, 1, 128) and big.supported(n, nnz, 1, 32)
    _lib.lib().dfgnn_set_block_mode(3)
    assert plan.algorithm(n, nnz, 1, 128, True) == 2 and plan.algorithm(n, nnz, 1, 128, False) == 0
    assert plan.algorithm(n, nnz, 1, 32, True) == 0 and big.algorithm(n, nnz, 1, 128, True) == 0
    _lib.lib().dfgnn_set_block_mode(2)
    gv = graphs.pascalvoc_like(batch=2).to(cuda)
    rp, ci, rws, vl, sm = preprocess_Hyper(gv)
    X = graphs.conv_inputs(gv.num_nodes(), 128, 1)
    N.gt_hyper_inference(rp, ci, rws, vl, sm, X.Q.to(cuda), X.K.to(cuda), X.V.to(cuda))
    assert _lib.last_kernel(0) == "dot_fwd_kernel"


@pytest.mark.parametrize("kind,dim", [("pattern", 128), ("pattern", 64), ("pattern-max", 128), ("ragged", 128),
                                      ("ragged", 64)])
def test_dense_tensor_core_forward(cuda, kind, dim):
    """dense_gt.cuh: masked 16-row tiles on the tensor cores (mma.sync TF32, 3xTF32 split) --
    out and attn_edge against the fp64 oracle, training and inference entry points."""
    _lib.lib().dfgnn_set_block_mode(3)
    g = _batch(kind)
    n = g.num_nodes()
    X = graphs.conv_inputs(n, dim, 23)
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(cuda))
    Q, K, V, dO = (t.to(cuda) for t in (X.Q, X.K, X.V, X.dO))
    out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    assert _lib.last_kernel(0) == "gt_dense_fwd_kernel"
    o64, a64, dQ, dK, dV = _oracle(g, X)
    assert_close("out", out, o64)
    assert_close("attn_edge", attn, a64)
    inf = N.gt_hyper_inference(row_ptr, col_ind, rows, val, smem, Q, K, V)[0]
    assert _lib.last_kernel(0) == "gt_dense_fwd_kernel"
    assert torch.equal(inf, out)
    # the backward (general or staged kernels) consumes the dense forward's attn_edge
    gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO)
    assert_close("grad_Q", gq, dQ)
    assert_close("grad_K", gk, dK)
    assert_close("grad_V", gv, dV)
    # automatic mode picks the dense kernels from the fill ratio of the blocks: tcgen05 (dense_tc.cu) at
    # f = 128 for training and inference, mma.sync at f = 64 for inference only
    _lib.lib().dfgnn_set_block_mode(0)
    plan = row_ptr._dfgnn_blocks
    dense = col_ind.numel() / plan.sum_sq_nodes >= plan.DENSE_MIN_FILL
    tc = dim == 128 and plan.prefers_dense_tc(col_ind.numel())  # dense AND enough edges per 128-row tile
    assert plan.algorithm(n, col_ind.numel(), 1, dim, True) == (3 if tc else 2 if dense else 0)
    assert plan.algorithm(n, col_ind.numel(), 1, dim, True, training=True) == (3 if tc else 0)
    inf2 = N.gt_hyper_inference(row_ptr, col_ind, rows, val, smem, Q, K, V)[0]
    assert _lib.last_kernel(0) == ("gt_dense_tc_fwd_kernel" if tc else "gt_dense_fwd_kernel" if dense else "dot_fwd_kernel")
    assert_close("automatic mode inference", inf2, o64)


def _batch_tc(kind):
    if kind == "wide":   # graphs of 130-256 nodes: two row tiles, two key halves, partial last slices
        return graphs.batched_graph(5, 200.0, 40.0, 130, 256, 40.0, 15.0, 1, None, 6, "wide")
    return _batch(kind)


@pytest.mark.parametrize("kind", ["pattern", "pattern-max", "ragged", "wide"])
def test_dense_tcgen05_forward(cuda, kind):
    """dense_tc.cu: per 128-row tile S = Q K^T and O = P V on tcgen05 (3xTF32, accumulators in tensor
    memory), softmax from tensor memory -- out and attn_edge against the fp64 oracle, training and
    inference entry points, many more graphs than SMs would be needed to wrap the persistent loop, so
    the batch is also run twice through a 2-CTA-sized plan below."""
    _lib.lib().dfgnn_set_block_mode(4)
    g = _batch_tc(kind)
    n = g.num_nodes()
    X = graphs.conv_inputs(n, 128, 29)
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(cuda))
    plan = row_ptr._dfgnn_blocks
    assert plan.ascending and plan.algorithm(n, col_ind.numel(), 1, 128, True, training=True) == 3
    Q, K, V, dO = (t.to(cuda) for t in (X.Q, X.K, X.V, X.dO))
    out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    assert _lib.last_kernel(0) == "gt_dense_tc_fwd_kernel"
    o64, a64, dQ, dK, dV = _oracle(g, X)
    assert_close("out", out, o64)
    assert_close("attn_edge", attn, a64)
    inf = N.gt_hyper_inference(row_ptr, col_ind, rows, val, smem, Q, K, V)[0]
    assert _lib.last_kernel(0) == "gt_dense_tc_fwd_kernel"
    assert torch.equal(inf, out)
    # backward: both sides on tcgen05 (dense P / dS tiles in between)
    gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO)
    assert _lib.last_kernel(1) == "gt_dense_tc_bwd_row_kernel" and _lib.last_kernel(2) == "gt_dense_tc_bwd_col_kernel"
    # the column-side kernel also runs from the general row-side kernel's packed CSR scratch
    bufs = N.gt_backward(row_ptr.clone(), col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO, _phases=1)
    plan_ = row_ptr._dfgnn_blocks
    nc_, sp_, si_ = plan_.col_sched
    gk_c, gv_c = torch.empty_like(K), torch.empty_like(V)
    rc = _lib.lib().dfgnn_gt_dense_tc_backward_col(
        plan_.n_blocks, plan_.blk_ptr.data_ptr(), plan_.max_nodes, n, col_ind.numel(), 1, 128, row_ptr.data_ptr(),
        plan_.adj_bits.data_ptr(), nc_, sp_.data_ptr(), si_.data_ptr(), Q.data_ptr(), dO.data_ptr(), bufs[3].data_ptr(),
        gk_c.data_ptr(), gv_c.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "dense_tc_backward_col")
    assert_close("grad_K from the CSR scratch", gk_c, dK)
    assert_close("grad_V from the CSR scratch", gv_c, dV)
    # grad_Q = sum_j p_ij (dA_ij - s_i) K_j cancels heavily and goes through two 3xTF32 products whose fp32
    # accumulation in the tensor core truncates: a few elements per million land up to 1.5x outside
    # 1e-4 / 1e-5 (tools/tc_error.py; the general fp32 kernel reaches 0.95x on the same inputs) -- the
    # full-size allowance (every element within 2x, at most max(1, 1e-6 of them) outside 1x) applies
    assert_close_bulk("grad_Q", gq[:, 0], torch.from_numpy(dQ)[:, 0])
    assert_close("grad_K", gk, dK)
    assert_close("grad_V", gv, dV)
    # other widths and weighted scores stay on the other kernels
    assert plan.algorithm(n, col_ind.numel(), 1, 64, True) != 3
    assert plan.algorithm(n, col_ind.numel(), 1, 128, False) != 3


def test_dense_tcgen05_persistent_loop_and_unsorted_fallback(cuda):
    """600 graphs on 148 persistent CTAs (every CTA walks several graphs: ring slots, tensor memory
    and barrier phases are reused across tiles); a CSR whose rows are not sorted by column falls
    back to the general kernels."""
    _lib.lib().dfgnn_set_block_mode(4)
    g = graphs.batched_graph(600, 60.0, 50.0, 1, 200, 20.0, 15.0, 0, None, 8, "many")
    n = g.num_nodes()
    X = graphs.conv_inputs(n, 128, 31)
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(cuda))
    Q, K, V, dO = (t.to(cuda) for t in (X.Q, X.K, X.V, X.dO))
    out = N.gt_hyper_inference(row_ptr, col_ind, rows, val, smem, Q, K, V)[0]
    assert _lib.last_kernel(0) == "gt_dense_tc_fwd_kernel"
    rp2 = row_ptr.clone()  # no plan attached: general kernels
    out2 = N.gt_hyper_inference(rp2, col_ind, rows, val, smem, Q, K, V)[0]
    assert _lib.last_kernel(0) == "dot_fwd_kernel"
    assert_close("dense tcgen05 vs general kernels", out, out2)
    # training forward + backward through the persistent loops (several tiles / items per CTA)
    o_t, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    assert _lib.last_kernel(0) == "gt_dense_tc_fwd_kernel"
    gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO)
    assert _lib.last_kernel(2) == "gt_dense_tc_bwd_col_kernel"
    from .test_fullsize_gpu import _gt_reference_chunked
    ro, ra, rq, rk, rv = _gt_reference_chunked(row_ptr, col_ind, Q, K, V, dO)  # fp64 on the device
    assert_close("persistent loop out", o_t[:, 0], ro)
    assert_close("persistent loop attn_edge", attn[0], ra)
    assert_close_bulk("persistent loop grad_Q", gq[:, 0], rq)  # see test_dense_tcgen05_forward
    assert_close("persistent loop grad_K", gk[:, 0], rk)
    assert_close("persistent loop grad_V", gv[:, 0], rv)
    # reverse the column order inside every row: still a valid CSR, no longer ascending
    rp = row_ptr.long()
    pos = torch.arange(col_ind.numel(), device=cuda)
    rowid = torch.repeat_interleave(torch.arange(n, device=cuda), rp[1:] - rp[:-1])
    rev = (rp[rowid] + rp[rowid + 1] - 1 - pos)
    ci_rev = col_ind[rev].contiguous()
    plan = formats.block_plan(g.batch_num_nodes(), row_ptr, ci_rev)
    assert not plan.ascending and plan.algorithm(n, ci_rev.numel(), 1, 128, True) != 3


def test_adjacency_bitmap_and_schedules_are_exact(cuda):
    """The formats of the dense tcgen05 kernels (integer work: bit-exact): adj_bits[r] has bit j set iff
    (r, first node of r's graph + j) is an edge; the work lists hold every graph / (graph, key tile) item
    exactly once."""
    g = _batch_tc("wide")
    row_ptr, col_ind, rows, val, smem = preprocess_Hyper(g.to(cuda))
    plan = row_ptr._dfgnn_blocks
    assert plan.ascending and plan.adj_bits is not None and plan.adj_bits.shape == (g.num_nodes(), 8)
    rp, ci = row_ptr.cpu().numpy(), col_ind.cpu().numpy()
    blk = plan.blk_ptr.cpu().numpy()
    want = np.zeros((g.num_nodes(), 8), dtype=np.uint32)
    for b in range(plan.n_blocks):
        for r in range(blk[b], blk[b + 1]):
            j = ci[rp[r]:rp[r + 1]] - blk[b]
            np.bitwise_or.at(want[r], j >> 5, (np.uint32(1) << (j & 31).astype(np.uint32)))
    assert np.array_equal(plan.adj_bits.cpu().numpy().view(np.uint32), want)
    assert sorted(plan.sched_idx.cpu().tolist()) == list(range(plan.n_blocks))
    bnn = g.batch_num_nodes().numpy()
    items = sorted([2 * b for b in range(plan.n_blocks)] + [2 * b + 1 for b in range(plan.n_blocks) if bnn[b] > 128])
    assert sorted(plan.col_sched[2].cpu().tolist()) == items


def test_dense_tcgen05_path_is_bit_reproducible(cuda):
    """No atomics anywhere on the dense tcgen05 path (the reference's backward accumulates with atomicAdd):
    two runs on the same inputs give identical bits, forward and backward."""
    _lib.lib().dfgnn_set_block_mode(4)
    g = _batch_tc("wide")
    n = g.num_nodes()
    X = graphs.conv_inputs(n, 128, 37)
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(cuda))
    Q, K, V, dO = (t.to(cuda) for t in (X.Q, X.K, X.V, X.dO))
    runs = []
    for _ in range(2):
        out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
        gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO)
        runs.append((out, attn, gq, gk, gv))
    assert _lib.last_kernel(0) == "gt_dense_tc_fwd_kernel" and _lib.last_kernel(1) == "gt_dense_tc_bwd_row_kernel"
    for a, b in zip(*runs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("kind", ["wide", "ragged"])
def test_dense_tcgen05_kernels_read_no_stale_tensor_memory(cuda, kind):
    """Tensor memory keeps its contents between kernels.  Poison all of it with NaN, then run forward and
    backward: a column read without having been written (padding keys of a wide graph, the second
    accumulator half) would turn up as NaN."""
    _lib.lib().dfgnn_set_block_mode(4)
    g = _batch_tc(kind)
    n = g.num_nodes()
    X = graphs.conv_inputs(n, 128, 41)
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(cuda))
    Q, K, V, dO = (t.to(cuda) for t in (X.Q, X.K, X.V, X.dO))
    st = torch.cuda.current_stream().cuda_stream
    o64, a64, dQ, dK, dV = _oracle(g, X)
    _lib.check(_lib.lib().dfgnn_tc_poison_tmem(st), "poison")
    out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    assert _lib.last_kernel(0) == "gt_dense_tc_fwd_kernel"
    _lib.check(_lib.lib().dfgnn_tc_poison_tmem(st), "poison")
    bufs = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO, _phases=1)
    _lib.check(_lib.lib().dfgnn_tc_poison_tmem(st), "poison")
    gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO,
                               _phases=2, _buffers=bufs)
    for t in (out, attn, gq, gk, gv):
        assert bool(torch.isfinite(t).all())
    assert_close("out", out, o64)
    assert_close("attn_edge", attn, a64)
    assert_close_bulk("grad_Q", gq[:, 0], torch.from_numpy(dQ)[:, 0])
    assert_close("grad_K", gk, dK)
    assert_close("grad_V", gv, dV)
