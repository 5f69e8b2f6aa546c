"""Generate golden vectors from the REFERENCE's own CUDA kernels (oracle/_ref,
compiled unchanged for sm_100a) on a B200:

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden
    cp gpurun_out/golden/*.npz tests/golden/

Each fixture stores only the graph/seed recipe plus the reference outputs; the
inputs are regenerated from the seeds (dfgnn_b200/graphs.py) by the tests.  The
reference holds no golden vectors of its own, so these
fixtures are what pins the CPU oracle (tests/test_oracle_cpu.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from dfgnn_b200 import graphs  # noqa: E402
from oracle import cpu_oracle as O  # noqa: E402
from oracle import ref_gpu  # noqa: E402

# name -> (graph recipe, dim, seed); small enough to commit (tens of KB each)
RECIPES = {
    "cora_d32": ("cora_like", dict(scale=0.12), 32, 21),
    "arxiv_d64": ("arxiv_like", dict(scale=0.002), 64, 22),
    "pattern_d128": ("pattern_like", dict(batch=2), 128, 23),
    "voc_d64": ("pascalvoc_like", dict(batch=1), 64, 24),
}


def build(name):
    fn, kw, dim, seed = RECIPES[name]
    g = getattr(graphs, fn)(**kw)
    src, dst = g.edges()
    n = g.num_nodes()
    rp, ci, rows, perm = O.coo_to_csr(src, dst, n)
    cp, ri, vi = O.csr_to_csc(rp, ci, n)
    X = graphs.conv_inputs(n, dim, seed)
    return g, dict(row_ptr=rp, col_ind=ci, rows=rows, col_ptr=cp, row_ind=ri, val_idx=vi), X


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    gt, gat = ref_gpu.fused_gtconv(), ref_gpu.fused_gatconv()
    assert gt is not None and gat is not None, "oracle/_ref is not built"
    dev = torch.device("cuda:0")
    for name in RECIPES:
        g, idx, X = build(name)
        d = {k: torch.from_numpy(v).to(dev) for k, v in idx.items()}
        val = torch.ones(len(idx["col_ind"]), device=dev)
        Q, K, V, dO = (t.to(dev).contiguous() for t in (X.Q, X.K, X.V, X.dO))
        ar, ac = X.attn_row.to(dev).contiguous(), X.attn_col.to(dev).contiguous()
        hs = ref_gpu.hyper_smem(d["row_ptr"])
        out, attn = gt.gt_hyper_forward(d["row_ptr"], d["col_ind"], d["rows"], val, d["col_ptr"],
                                        d["row_ind"], d["val_idx"], hs, Q, K, V)
        gq, gk, gv = gt.gt_backward(d["row_ptr"], d["col_ind"], d["rows"], val, d["col_ptr"],
                                    d["row_ind"], d["val_idx"], hs, Q, K, V, attn, dO)
        til = gt.gt_tiling_inference(d["row_ptr"], d["col_ind"], val, 128, Q, K, V)[0]
        g_out, emax, esum, emask = gat.gat_forward(ar, ac, d["row_ptr"], d["col_ind"], 0.2, V, 0.0)
        gf, gr, gc = gat.gat_backward(0.2, 0.0, d["row_ptr"], d["col_ind"], d["col_ptr"],
                                      d["row_ind"], d["val_idx"], emax, esum, emask, V, ar, ac, dO)
        torch.cuda.synchronize()
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            recipe=np.array(repr(RECIPES[name])), sha256=np.array(g.sha256()),
            gt_out=out.cpu().numpy(), gt_attn=attn.cpu().numpy(), gt_tiling_out=til.cpu().numpy(),
            gt_dQ=gq.cpu().numpy(), gt_dK=gk.cpu().numpy(), gt_dV=gv.cpu().numpy(),
            gat_out=g_out.cpu().numpy(), gat_emax=emax.cpu().numpy(), gat_esum=esum.cpu().numpy(),
            gat_dfeat=gf.cpu().numpy(), gat_drow=gr.cpu().numpy(), gat_dcol=gc.cpu().numpy())
        print(name, g, "hyper_smem", hs, flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
