#!/usr/bin/env python
"""Per-kernel timing of the conv kernels on the bench workloads (developer tool).

    python tools/kbench.py [--workloads arxiv-gat,pattern-gt,voc-gt] [--iters 10] [--tag NAME]

For each workload: fwd and bwd through dfgnn_b200.operators._native on resident
operands, L2 flushed between iterations, per-kernel CUDA durations collected with CUPTI
(torch.profiler) so the row-side and column-side backward kernels are separated.
`DFGNN_B200_LIB=/path/to/variant.so` selects a library variant (see dfgnn_b200/_lib.py).
Prints one line per kernel and a JSON summary line; also a checksum of every output so
variants can be compared for agreement.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from torch.profiler import ProfilerActivity, profile

    import bench as B
    from dfgnn_b200 import graphs
    from dfgnn_b200.layers import preprocess_gat_fw_bw, preprocess_Hyper_fw_bw
    from dfgnn_b200.operators import _native as N

    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="arxiv-gat,pattern-gt,voc-gt")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tag", default=os.environ.get("DFGNN_B200_LIB", "default"))
    ap.add_argument("--reddit-scale", type=float, default=1.0)
    ap.add_argument("--dim", type=int, default=0, help="override the feature width of the workloads")
    ap.add_argument("--conv", default="", help="override the conv (gt | gat)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    peak, _ = B.hbm_peak()
    summary = {"tag": args.tag}

    for name in args.workloads.split(","):
        if name.startswith("rand:"):  # rand:<mean degree> -- 200k-node full graph, log-normal degrees
            deg = float(name.split(":")[1])
            conv, dim, fn, kw, fmt, cfg = "gat", 64, None, {}, "softmax", -1
            B.SEEDS[name] = 77
        else:
            conv, dim, fn, kw, fmt, cfg = B.WORKLOADS[name]
        dim = args.dim or dim
        conv = args.conv or conv
        if name.startswith("rand:"):
            g = graphs.full_graph(200000, int(200000 * deg), deg, 0.4 * deg, 1, int(4 * deg), 77, name)
        elif name == "reddit-gt" and args.reddit_scale != 1.0:
            g = graphs.reddit_like(args.reddit_scale)
        else:
            g = B.build_graph(name)
        n, e = g.num_nodes(), g.num_edges()
        gd = g.to(dev)
        X = graphs.conv_inputs(n, dim, B.SEEDS[name])
        if conv == "gt":
            A, rows, rp, ci, val, cp, ri, vi, smem = preprocess_Hyper_fw_bw(gd)
            Q, K, V, dO = (t.to(dev) for t in (X.Q, X.K, X.V, X.dO))

            def step():
                out, attn = N.gt_hyper_forward(rp, ci, rows, val, cp, ri, vi, smem, Q, K, V)
                return (out, attn) + tuple(N.gt_backward(rp, ci, rows, val, cp, ri, vi, smem, Q, K, V, attn, dO))
        else:
            rp, ci, cp, ri, vi = preprocess_gat_fw_bw(gd)
            ar, ac, F, dO = (t.to(dev) for t in (X.attn_row, X.attn_col, X.V, X.dO))

            def step():
                out, emax, esum, emask = N.gat_forward(ar, ac, rp, ci, 0.2, F, 0.0)
                return (out, esum) + tuple(N.gat_backward(0.2, 0.0, rp, ci, cp, ri, vi, emax, esum, emask,
                                                          F, ar, ac, dO))
        for _ in range(3):
            res = step()
        torch.cuda.synchronize()
        sums = [float(t.double().abs().sum()) for t in res]
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(args.iters):
                flush.fill_(1.0)
                step()
            torch.cuda.synchronize()
        # whole step (all launches) with CUDA events, L2 flushed before each step
        evs = []
        for _ in range(args.iters):
            flush.fill_(1.0)
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b_.record()
            evs.append((a, b_))
        torch.cuda.synchronize()
        step_us = sorted(a.elapsed_time(b_) * 1e3 for a, b_ in evs)[len(evs) // 2]
        rows_out = {}
        for ev in prof.key_averages():
            if "dfgnn" not in ev.key:
                continue
            short = ev.key.split("dfgnn::")[1].split("<")[0] if "dfgnn::" in ev.key else ev.key
            t = getattr(ev, "device_time_total", None)
            if t is None:
                t = ev.cuda_time_total
            rows_out[short] = rows_out.get(short, 0.0) + t / max(ev.count, 1)
        tot = sum(rows_out.values())
        fwd_k = "gat_fwd_staged_kernel" if "gat_fwd_staged_kernel" in rows_out else (
            "gat_fwd_kernel" if conv == "gat" else "dot_fwd_kernel")
        fb = B.alg_bytes(conv, "fwd", n, e, dim)
        sb = B.alg_bytes(conv, "fwd+bwd", n, e, dim)
        print(f"[{args.tag}] {name}: N={n} E={e} d={dim}  step(events, median) {step_us:.1f} us  kernels total {tot:.1f} us "
              f"(step frac {sb / tot / 1e3 / peak:.3f})  fwd frac {fb / rows_out.get(fwd_k, 1e9) / 1e3 / peak:.3f}")
        for k, v in rows_out.items():
            print(f"    {k:24s} {v:9.1f} us")
        print("    checksums", " ".join(f"{s:.6e}" for s in sums), flush=True)
        summary[name] = {"total_us": tot, **rows_out}
    print("KBENCH " + json.dumps(summary), flush=True)


if __name__ == "__main__":
    main()
