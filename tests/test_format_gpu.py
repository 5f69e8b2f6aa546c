"""Format construction on the GPU: bit-exact against the CPU oracle and scipy
(integer work; SURVEY.md 8c defines CSR = stable sort by row, CSC = stable sort
by column with val_idx = CSC position -> CSR position)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from dfgnn_b200 import formats, graphs
from dfgnn_b200.layers import (preprocess_CSR, preprocess_gat_fw_bw, preprocess_Hyper,
                               preprocess_Hyper_fw_bw, preprocess_softmax)
from oracle import cpu_oracle as O

from .helpers import random_graph

pytestmark = pytest.mark.gpu

CASES = {
    "cora": lambda: graphs.cora_like(),
    "arxiv-small": lambda: graphs.arxiv_like(0.05),
    "pattern": lambda: graphs.pattern_like(batch=16),
    "voc": lambda: graphs.pascalvoc_like(batch=8),
    "holes": lambda: random_graph(1000, 5, 3, max_deg=900, empty_frac=0.4),
    "one-edge": lambda: graphs.Graph(torch.tensor([2]), torch.tensor([0]), 3),
    "no-edges": lambda: graphs.Graph(torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), 9),
}


def _eq(name, got, want):
    got = got.cpu().numpy()
    assert got.dtype == want.dtype, f"{name}: dtype {got.dtype} vs {want.dtype}"
    assert np.array_equal(got, want), f"{name} differs"


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("shuffle", [False, True])
def test_csr_csc_bit_exact(cuda, name, shuffle):
    g = CASES[name]()
    src, dst = g.edges()
    n = g.num_nodes()
    if shuffle and src.numel() > 1:  # arbitrary input order: the sort must be STABLE by row
        p = torch.randperm(src.numel(), generator=torch.Generator().manual_seed(1))
        src, dst = src[p], dst[p]
    rp, ci, rows, perm = O.coo_to_csr(src, dst, n)
    cp, ri, vi = O.csr_to_csc(rp, ci, n)
    d_rp, d_ci, d_rows, d_perm, d_val = formats.coo_to_csr(src.to(cuda), dst.to(cuda), n)
    _eq("row_ptr", d_rp, rp)
    _eq("col_ind", d_ci, ci)
    _eq("rows", d_rows, rows)
    _eq("perm", d_perm, perm)
    assert d_val.dtype == torch.float32 and bool((d_val == 1).all())
    d_cp, d_ri, d_vi = formats.csr_to_csc(d_rp, d_ci)
    _eq("col_ptr", d_cp, cp)
    _eq("row_ind", d_ri, ri)
    _eq("val_idx", d_vi, vi)
    # the same with the expanded row ids handed over (row_ind by gather instead of search)
    for a, b in zip(formats.csr_to_csc(d_rp, d_ci, None, d_rows), (d_cp, d_ri, d_vi)):
        assert torch.equal(a, b)
    # second oracle: scipy (train_gatconv.py:119-136 builds `permute` exactly like this)
    if len(ci):
        A = sp.csr_matrix((np.arange(len(ci), dtype=np.int32), ci, rp), shape=(n, n)).tocsc()
        _eq("col_ptr/scipy", d_cp, A.indptr.astype(np.int32))
        _eq("row_ind/scipy", d_ri, A.indices.astype(np.int32))
        _eq("val_idx/scipy", d_vi, A.data.astype(np.int32))


def test_preprocess_tuples_match_reference_layout(cuda):
    g = graphs.pattern_like(batch=4).to(cuda)
    src, dst = g.edges()
    n = g.num_nodes()
    rp, ci, rows, _ = O.coo_to_csr(src, dst, n)
    cp, ri, vi = O.csr_to_csc(rp, ci, n)
    row_ptr, col_ind, val, smem = preprocess_CSR(g)          # layers/util.py:79
    assert smem == 128
    _eq("row_ptr", row_ptr, rp); _eq("col_ind", col_ind, ci)
    row_ptr, col_ind, rws, val, smem = preprocess_Hyper(g)   # layers/util.py:100
    assert smem == 1024
    _eq("rows", rws, rows)
    row_ptr, col_ind, rws, val, smem = preprocess_softmax(g) # layers/util.py:162
    assert smem == 128 and val.dtype == torch.float32
    A, rws, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g)
    assert smem == 1024 and A.shape == (n, n)
    _eq("col_ptr", col_ptr, cp); _eq("row_ind", row_ind, ri); _eq("val_idx", val_idx, vi)
    assert preprocess_Hyper_fw_bw(g, fused=False)[1:] == (None,) * 8
    row_ptr, col_ind, col_ptr, row_ind, permute = preprocess_gat_fw_bw(g)
    _eq("permute", permute, vi)


def test_format_rejects_bad_input(cuda):
    with pytest.raises(RuntimeError, match="outside"):
        formats.coo_to_csr(torch.tensor([0, 5], device=cuda), torch.tensor([1, 1], device=cuda), 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        formats.coo_to_csr(torch.tensor([0]), torch.tensor([0]), 3)
    # a CSR whose column ids do not fit the stated column count (a shard handed the wrong n_cols)
    rp = torch.tensor([0, 2, 3], dtype=torch.int32, device=cuda)
    ci = torch.tensor([0, 9, 1], dtype=torch.int32, device=cuda)
    with pytest.raises(RuntimeError, match="column index outside"):
        formats.csr_to_csc(rp, ci, 4)
    with pytest.raises(RuntimeError, match="column index outside"):
        formats.csr_to_csc(rp, torch.tensor([0, -1, 1], dtype=torch.int32, device=cuda), 4)
    cp, ri, vi = formats.csr_to_csc(rp, ci, 10)   # fine once the column count is right
    assert cp.tolist() == [0, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3] and ri.tolist() == [0, 1, 0] and vi.tolist() == [0, 2, 1]


def test_full_size_arxiv_and_pattern_properties(cuda):
    """BASELINE.json sizes: checked through size-independent properties."""
    for g in (graphs.arxiv_like(), graphs.pattern_like()):
        src, dst = g.edges()
        n = g.num_nodes()
        rp, ci, rows, perm, val = formats.coo_to_csr(src.to(cuda), dst.to(cuda), n)
        cp, ri, vi = formats.csr_to_csc(rp, ci)
        E = src.numel()
        assert int(rp[0]) == 0 and int(rp[-1]) == E and int(cp[-1]) == E
        assert bool((rp[1:] >= rp[:-1]).all()) and bool((cp[1:] >= cp[:-1]).all())
        assert bool((rows[1:] >= rows[:-1]).all())                     # sortedness
        assert torch.equal(torch.sort(vi.long()).values, torch.arange(E, device=cuda))  # permutation
        assert torch.equal(ci.long()[vi.long()], torch.repeat_interleave(
            torch.arange(n, device=cuda), (cp[1:] - cp[:-1]).long()))   # CSC entry p sits in column col_ind[val_idx[p]]
        assert torch.equal(ri, rows[vi.long()])                         # and in row rows[val_idx[p]]
        assert torch.equal(src.to(cuda)[perm.long()], rows.long())      # perm maps back to the COO
