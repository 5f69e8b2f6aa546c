"""Forward timings on the PATTERN-shaped batch: general vs block vs dense kernels, training and
inference entry points (developer tool).  python tools/time_pattern.py [dim]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfgnn_b200 import _lib, graphs
from dfgnn_b200.layers import preprocess_Hyper_fw_bw
from dfgnn_b200.operators import _native as N

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
g = graphs.pattern_like()
n = g.num_nodes()
A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(dev))
X = graphs.conv_inputs(n, dim, 3)
Q, K, V, dO = (t.to(dev) for t in (X.Q, X.K, X.V, X.dO))
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timeit(fn, it=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(it):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


print("lib", _lib.LIB_PATH)
for mode, name in ((1, "general"), (2, "staged"), (3, "dense"), (4, "dense-tc")):
    _lib.lib().dfgnn_set_block_mode(mode)
    t_tr = timeit(lambda: N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V))
    k = _lib.last_kernel(0)
    t_inf = timeit(lambda: N.gt_hyper_inference(row_ptr, col_ind, rows, val, smem, Q, K, V))
    out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
    t_bwd = timeit(lambda: N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO))
    print(f"{name:8s} {k:24s} train fwd {t_tr:.4f} ms   inference {t_inf:.4f} ms   backward {t_bwd:.4f} ms ({_lib.last_kernel(1)})")
