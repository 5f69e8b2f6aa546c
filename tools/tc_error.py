"""Error statistics of the dense tcgen05 GT kernels against an fp64 restatement on the device, next to the
general fp32 kernels on the same inputs (developer tool).  python tools/tc_error.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfgnn_b200 import _lib, graphs
from dfgnn_b200.layers import preprocess_Hyper_fw_bw
from dfgnn_b200.operators import _native as N
from tests.test_fullsize_gpu import _gt_reference_chunked

dev = torch.device("cuda:0")
batches = {
    "pattern(256 graphs)": graphs.pattern_like(batch=256),
    "wide": graphs.batched_graph(64, 200.0, 40.0, 130, 256, 40.0, 15.0, 1, None, 6, "wide"),
    "many": graphs.batched_graph(600, 60.0, 50.0, 1, 200, 20.0, 15.0, 0, None, 8, "many"),
}


def stats(name, got, want):
    err = (got.double() - want).abs()
    tol = 1e-5 + 1e-4 * want.abs()
    ratio = err / tol
    return f"{name}: worst {float(ratio.max()):5.2f}x  outside {int((ratio > 1).sum()):4d}/{err.numel()}  rms {float((ratio ** 2).mean().sqrt()):.4f}"


for label, g in batches.items():
    n = g.num_nodes()
    A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(dev))
    X = graphs.conv_inputs(n, 128, 3)
    Q, K, V, dO = (t.to(dev) for t in (X.Q, X.K, X.V, X.dO))
    ro, ra, rq, rk, rv = _gt_reference_chunked(row_ptr, col_ind, Q, K, V, dO)
    for mode, mname in ((1, "general fp32"), (4, "tcgen05 3xTF32")):
        _lib.lib().dfgnn_set_block_mode(mode)
        out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
        gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO)
        print(f"[{label}] {mname} ({_lib.last_kernel(0)}, {_lib.last_kernel(1)}, {_lib.last_kernel(2)})")
        for nm, a, b in (("out   ", out[:, 0], ro), ("attn  ", attn[0], ra), ("grad_Q", gq[:, 0], rq), ("grad_K", gk[:, 0], rk),
                         ("grad_V", gv[:, 0], rv)):
            print("   ", stats(nm, a, b))
