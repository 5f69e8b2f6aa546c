"""Timing of the tcgen05 dense GT kernels on the PATTERN-shaped batch (developer tool).
python tools/time_tc.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfgnn_b200 import _lib, graphs
from dfgnn_b200.layers import preprocess_Hyper_fw_bw
from dfgnn_b200.operators import _native as N

it = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
g = graphs.pattern_like()
n = g.num_nodes()
A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(dev))
X = graphs.conv_inputs(n, 128, 3)
Q, K, V, dO = (t.to(dev) for t in (X.Q, X.K, X.V, X.dO))
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
_lib.lib().dfgnn_set_block_mode(4)


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(it):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


t_tr = timeit(lambda: N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V))
k0 = _lib.last_kernel(0)
t_inf = timeit(lambda: N.gt_hyper_inference(row_ptr, col_ind, rows, val, smem, Q, K, V))
out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
t_bwd = timeit(lambda: N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO))
print(f"{k0}: train fwd {t_tr*1e3:.1f} us  inference {t_inf*1e3:.1f} us   backward {t_bwd*1e3:.1f} us "
      f"({_lib.last_kernel(1)}, {_lib.last_kernel(2)})")

if os.environ.get("DFGNN_B200_LIB"):  # profiling build (-DDFGNN_TC_PROF): per-role cycle counters of block 0
    import ctypes
    h = ctypes.CDLL(os.environ["DFGNN_B200_LIB"])
    if hasattr(h, "dfgnn_tc_prof_read"):
        buf = (ctypes.c_ulonglong * 32)()
        names = {0: "soft wait mask", 1: "soft wait S", 2: "soft wait P slot", 3: "soft wait O", 4: "soft pass1",
                 5: "soft pass2 (incl. slot waits)", 6: "soft pass3", 7: "soft epilogue", 8: "mma wait s_free",
                 9: "mma wait full_b (1)", 10: "mma wait o_free", 11: "mma wait full_b (2)", 12: "mma wait full_a",
                 16: "loader group 0 wait empty (1)", 17: "loader group 0 wait empty (2)", 18: "loader group 0 load latency (1)",
                 19: "loader group 0 load latency (2)", 24: "total softmax thread (group 0)", 26: "total mma thread", 27: "total loader thread"}
        for label, fn in (("training forward, L2 flushed", lambda: N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)),
                          ("inference, L2 flushed", lambda: N.gt_hyper_inference(row_ptr, col_ind, rows, val, smem, Q, K, V))):
            flush.fill_(1.0)
            h.dfgnn_tc_prof_read(buf, 1)
            fn()
            h.dfgnn_tc_prof_read(buf, 1)
            print(label, "(block 0, one thread per role)")
            for k in sorted(names):
                print(f"  {names[k]:36s} {buf[k] / 1965.0:9.1f} us")
