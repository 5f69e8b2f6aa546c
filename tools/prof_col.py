import os, sys, ctypes
sys.path.insert(0, "/root/repo")
import torch
from dfgnn_b200 import _lib, graphs
from dfgnn_b200.layers import preprocess_Hyper_fw_bw
from dfgnn_b200.operators import _native as N
dev = torch.device("cuda:0")
g = graphs.pattern_like(); n = g.num_nodes()
A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g.to(dev))
X = graphs.conv_inputs(n, 128, 3)
Q, K, V, dO = (t.to(dev) for t in (X.Q, X.K, X.V, X.dO))
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
_lib.lib().dfgnn_set_block_mode(4)
out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
bufs = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO, _phases=1)
h = ctypes.CDLL(os.environ["DFGNN_B200_LIB"]); buf = (ctypes.c_ulonglong * 32)()
for rep in range(2):
    flush.fill_(1.0); h.dfgnn_tc_prof_read(buf, 1)
    N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, Q, K, V, attn, dO, _phases=2, _buffers=bufs)
    h.dfgnn_tc_prof_read(buf, 1)
names = {0: "worker wait empty", 1: "worker wait acc_full", 7: "worker epilogue", 8: "mma wait acc_free", 9: "mma wait full_b", 12: "mma wait full_a",
         14: "mma issue + commit", 16: "loader wait empty", 24: "total worker", 26: "total mma", 27: "total loader"}
for k in sorted(names): print(f"  {names[k]:28s} {buf[k]/1965.0:9.1f} us")
