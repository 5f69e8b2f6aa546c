#!/usr/bin/env python
"""bench.py -- fused attention-conv forward+backward throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" is one fused conv forward + backward over one synthetic graph of the
named workload (BASELINE.json configs; default = configs[1], GAT d=64 on the
arxiv-shaped full graph).  Prints ONE JSON line (rank 0).

  value     edges*dim per second, inputs resident in HBM, CUDA-event timed per step,
            L2 flushed between steps (a 256 MB write), max over ranks.
  e2e       same metric through the public operator (dfgnn_b200.operators, i.e. the
            reference's autograd-Function API) with HOST operands: every step copies the
            node features / logits / upstream gradient from pinned host memory and reads
            the output and the gradients back.  The graph index (CSR/CSC) is built once
            and stays resident, like the reference's `params = preprocess_func(g)`.
  roofline  the forward kernel: algorithmic bytes (SURVEY.md 8d gather model) / its
            CUDA-event duration inside the timed steps, against MEASURED_PEAKS.json.
  cpu_baseline   the CPU oracle (oracle/dfgnn_oracle.c, OpenMP) on the same workload.
  gpu_reference  the reference's own CUDA kernels (oracle/_ref, sm_100a) timed the same way.

`--impl reference` times the reference's CPU path: dgl / PyG are not installable
offline, so it is the oracle port (kind "port") on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fused_conv_fwd_bwd_edges_x_dim_per_s"
UNIT = "edges*dim/s"

WORKLOADS = {
    # name: (conv, dim, graph fn, kwargs, format, BASELINE.json config index)
    "arxiv-gat": ("gat", 64, "arxiv_like", {}, "softmax", 1),
    "pattern-gt": ("gt", 128, "pattern_like", {"batch": 1024}, "hyper", 2),
    "reddit-gt": ("gt", 128, "reddit_like", {}, "tiling", 3),
    "voc-gt": ("gt", 128, "pascalvoc_like", {"batch": 1024}, "hyper", 4),
    "cora-gt": ("gt", 128, "cora_like", {}, "hyper", 0),
}
SEEDS = {"arxiv-gat": 1002, "pattern-gt": 1003, "reddit-gt": 1004, "voc-gt": 1005, "cora-gt": 1001}


def alg_bytes(conv: str, phase: str, n: int, e: int, d: int) -> float:
    """Algorithmic bytes of SURVEY.md 8(d) (gather model, fp32, h = 1)."""
    if conv == "gt":
        fwd = 8.0 * e * d + 8.0 * n * d + 4.0 * e + 4.0 * (n + 1)
        both = 24.0 * e * d + 24.0 * n * d + 36.0 * e + 12.0 * (n + 1)
    else:
        fwd = 4.0 * e * d + 4.0 * n * d + 8.0 * e + 8.0 * n
        both = 16.0 * e * d + 12.0 * n * d + 52.0 * e
    return {"fwd": fwd, "fwd+bwd": both, "bwd": both - fwd}[phase]


def kernel_alg_bytes(conv: str, n: int, e: int, d: int) -> dict:
    """Per-kernel split of the gather model: every edge fetches the neighbour
    rows the kernel consumes, the row operand is read once per row, every index / edge scalar once."""
    if conv == "gt":
        return {"fwd": 8.0 * e * d + 8.0 * n * d + 8.0 * e + 4.0 * n,            # K,V rows; Q, out; col_ind, attn_edge
                "bwd_row": 8.0 * e * d + 8.0 * n * d + 16.0 * e + 4.0 * n,      # V,K rows; dO, dQ; col_ind, attn, {dS,p}
                "bwd_col": 8.0 * e * d + 8.0 * n * d + 16.0 * e + 4.0 * n}      # dO,Q rows; dK, dV; row_ind, val_idx, {dS,p}
    return {"fwd": 4.0 * e * d + 4.0 * n * d + 8.0 * e + 12.0 * n,                # feat rows; out; col_ind, attn_col[j]
            "bwd_row": 4.0 * e * d + 4.0 * n * d + 16.0 * e + 16.0 * n,          # feat rows; dO; col_ind, attn_col[j], {de,p}
            "bwd_col": 4.0 * e * d + 4.0 * n * d + 16.0 * e + 8.0 * n}           # dO rows; grad_feat; row_ind, permute, {de,p}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle samples during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        sm, mx, reasons = [], [], set()
        with open(self.tmp.name) as fh:
            for line in fh:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def build_graph(name: str, seed_offset: int = 0):
    from dfgnn_b200 import graphs
    conv, dim, fn, kw, fmt, cfg = WORKLOADS[name]
    return getattr(graphs, fn)(seed=SEEDS[name] + seed_offset, **kw)


# ----------------------------------------------------------------------------- #
# reference arm: the CPU path (oracle port) on the host cores                    #
# ----------------------------------------------------------------------------- #

def cpu_step_factory(name: str, g):
    import numpy as np
    from dfgnn_b200 import graphs
    from oracle import cpu_oracle as O
    conv, dim, *_ = WORKLOADS[name]
    src, dst = g.edges()
    n = g.num_nodes()
    rp, ci, rows, perm = O.coo_to_csr(src, dst, n)
    cp, ri, vi = O.csr_to_csc(rp, ci, n)
    X = graphs.conv_inputs(n, dim, SEEDS[name])
    Q, K, V, dO = (np.ascontiguousarray(t.numpy()) for t in (X.Q, X.K, X.V, X.dO))
    ar, ac = X.attn_row.numpy(), X.attn_col.numpy()
    if conv == "gt":
        def step():
            out, attn = O.gt_forward(rp, ci, None, Q, K, V)
            return O.gt_backward(rp, ci, cp, ri, vi, Q, K, V, attn, dO)
    else:
        def step():
            out, emax, esum = O.gat_forward(ar, ac, rp, ci, 0.2, V)
            return O.gat_backward(0.2, 0.0, rp, ci, cp, ri, vi, emax, esum, None, V, ar, ac, dO)
    return step, n, len(ci), dim


def time_cpu(name: str, g, steps: int, warmup: int):
    step, n, e, dim = cpu_step_factory(name, g)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return e * dim / dt, dt, n, e, dim


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and
    # may use all host cores (libgomp reads the variable when the oracle library is loaded)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    name = args.workload
    # bounded sample: the big workloads are shrunk so that K steps finish in minutes
    scale = {"reddit-gt": 0.05}.get(name, 1.0)
    if name == "reddit-gt":
        from dfgnn_b200 import graphs
        g = graphs.reddit_like(scale)
        sample = f"reddit-shaped graph at scale {scale} (N={g.num_nodes()}, E={g.num_edges()})"
    else:
        g = build_graph(name)
        sample = f"full workload (N={g.num_nodes()}, E={g.num_edges()})"
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    val, dt, n, e, dim = time_cpu(name, g, steps, warmup)
    cores = os.cpu_count() or 1
    conv, _, _, _, fmt, cfg = WORKLOADS[name]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "conv": conv, "dim": dim, "format": fmt, "nodes": n, "edges": e,
                   "baseline_config_index": cfg},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "oracle/dfgnn_oracle.c (OpenMP); the reference's DGL-sparse/PyG CPU "
                                 "path cannot be installed offline"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- #
# our arm                                                                        #
# ----------------------------------------------------------------------------- #

def run_ours(args):
    import torch
    import torch.distributed as dist

    from dfgnn_b200 import _lib, graphs
    from dfgnn_b200 import dist as ddist
    from dfgnn_b200.layers import preprocess_gat_fw_bw, preprocess_Hyper_fw_bw
    from dfgnn_b200.operators import GATConvFuse, GTConvFuse_hyper
    from dfgnn_b200.operators import _native as N

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def measure(mode: str, light: bool):
        """One measurement in partition mode `mode` ("weak" | "strong" | "auto"); `light` skips
        the e2e, per-kernel and baseline legs (used for the secondary row-partition numbers)."""
        name = args.workload
        conv, dim, fn, kw, fmt, cfg = WORKLOADS[name]
        batched = "batch" in kw
        weak = world > 1 and ((batched and mode != "strong") or mode == "weak")
        # ---- partition ------------------------------------------------
        if weak:
            # data-parallel: every rank owns its own batch of graphs, no collective in the conv
            g_full = build_graph(name, seed_offset=1000 * rank)
            part = ddist.make_partition(g_full, 1, 0)
            part.describe = (f"{kw['batch']} graphs per GPU" if batched else "one graph per GPU") + \
                f" on {world} GPUs (weak scaling), no collective"
        else:
            g_full = build_graph(name)
            part = ddist.make_partition(g_full, world, rank)
        n_total, e_total = g_full.num_nodes(), g_full.num_edges()
        g = part.local_graph.to(dev)
        n_rows, n_cols, e_local = part.n_rows, part.n_cols, part.local_graph.num_edges()

        X = graphs.conv_inputs(n_total, dim, SEEDS[name])
        rows_sl = part.row_slice
        pin = lambda t: t.contiguous().pin_memory()
        if conv == "gt":
            h_in = {"Q": pin(X.Q[rows_sl]), "K": pin(X.K[part.col_owned]), "V": pin(X.V[part.col_owned]),
                    "dO": pin(X.dO[rows_sl])}
        else:
            h_in = {"ar": pin(X.attn_row[rows_sl]), "ac": pin(X.attn_col[part.col_owned]),
                    "F": pin(X.V[part.col_owned]), "dO": pin(X.dO[rows_sl])}
        d_in = {k: v.to(dev) for k, v in h_in.items()}

        # resident index formats, built once by the CUDA format kernels
        if conv == "gt":
            A, rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem = preprocess_Hyper_fw_bw(g)
        else:
            row_ptr, col_ind, col_ptr, row_ind, val_idx = preprocess_gat_fw_bw(g)
            rows = val = None
            smem = 128
        torch.cuda.synchronize()

        # format construction timed separately, like the reference's own
        # `only_preprocess` loop (train_batch_graph_timing.py:115-143): COO -> CSR (+rows, val) -> CSC
        fmt_ms = []
        for _ in range(5):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if conv == "gt":
                preprocess_Hyper_fw_bw(g)
            else:
                preprocess_gat_fw_bw(g)
            b_.record()
            b_.synchronize()
            fmt_ms.append(a.elapsed_time(b_))
        fmt_ms = sorted(fmt_ms)[len(fmt_ms) // 2]

        halo = ddist.HaloExchange(part, dev, world)
        flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

        def step_device(rec=None):
            """fwd + bwd on resident operands; returns the tensors a caller would keep."""
            if conv == "gt":
                K, V = halo.gather_pair(d_in["K"], d_in["V"], rec)
                if rec is not None:
                    rec["f0"].record()
                out, attn = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx,
                                               smem, d_in["Q"], K, V)
                if rec is not None:
                    rec["f1"].record()
                gq, gk, gv = N.gt_backward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem,
                                           d_in["Q"], K, V, attn, d_in["dO"])
                gk, gv = halo.reduce_pair(gk, gv, rec)
                return out, gq, gk, gv
            F, ac = halo.gather_pair(d_in["F"], d_in["ac"], rec)
            if rec is not None:
                rec["f0"].record()
            out, emax, esum, emask = N.gat_forward(d_in["ar"], ac, row_ptr, col_ind, 0.2, F, 0.0)
            if rec is not None:
                rec["f1"].record()
            gf, gr, gc = N.gat_backward(0.2, 0.0, row_ptr, col_ind, col_ptr, row_ind, val_idx, emax, esum,
                                        emask, F, d_in["ar"], ac, d_in["dO"])
            gf, gc = halo.reduce_pair(gf, gc, rec)
            return out, gf, gr, gc

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        ev = lambda: torch.cuda.Event(enable_timing=True)
        for _ in range(max(args.warmup, 3)):
            step_device()
        barrier()

        # Single GPU: the step (6 launches + output allocations) is captured once in a CUDA graph and
        # replayed, so the timed region holds the kernels and not the Python launch path.  With
        # collectives in the step (N > 1 row partition) it runs eagerly.
        graph = None
        if not halo.active and not args.no_graph:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step_device()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_out = step_device()
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()

        clocks = ClockSampler(local) if rank == 0 else None
        n_launch0 = _lib.launch_count()
        launches_per_step = None
        recs = []
        barrier()
        for _ in range(args.steps):
            flush.fill_(1.0)  # L2 flush between timed steps (not timed)
            rec = {k: ev() for k in ("s", "f0", "f1", "e", "ag0", "ag1", "rs0", "rs1")}
            rec["s"].record()
            if graph is not None:
                graph.replay()
            else:
                step_device(rec)
            rec["e"].record()
            recs.append(rec)
        barrier()
        launches = _lib.launch_count() - n_launch0
        if graph is not None:
            # graph replays do not pass through the library's launch counter: count one eager step
            n0 = _lib.launch_count()
            step_device()
            launches = (_lib.launch_count() - n0) * args.steps
        # every kernel of the step on its own (roofline leg): same flush protocol, eager launches,
        # CUDA events around the single library call that launches it
        kern_ms = {"fwd": [], "bwd_row": [], "bwd_col": []}
        if not halo.active and not light:
            if conv == "gt":
                out0, attn0 = N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem,
                                                 d_in["Q"], d_in["K"], d_in["V"])
                bargs = (row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx, smem, d_in["Q"], d_in["K"],
                         d_in["V"], attn0, d_in["dO"])
                bufs = N.gt_backward(*bargs, _phases=1)
                calls = {"fwd": lambda: N.gt_hyper_forward(row_ptr, col_ind, rows, val, col_ptr, row_ind, val_idx,
                                                           smem, d_in["Q"], d_in["K"], d_in["V"]),
                         "bwd_row": lambda: N.gt_backward(*bargs, _phases=1, _buffers=bufs),
                         "bwd_col": lambda: N.gt_backward(*bargs, _phases=2, _buffers=bufs)}
            else:
                o0, emax0, esum0, emask0 = N.gat_forward(d_in["ar"], d_in["ac"], row_ptr, col_ind, 0.2, d_in["F"], 0.0)
                bargs = (0.2, 0.0, row_ptr, col_ind, col_ptr, row_ind, val_idx, emax0, esum0, emask0, d_in["F"],
                         d_in["ar"], d_in["ac"], d_in["dO"])
                bufs = N.gat_backward(*bargs, _phases=1)
                calls = {"fwd": lambda: N.gat_forward(d_in["ar"], d_in["ac"], row_ptr, col_ind, 0.2, d_in["F"], 0.0),
                         "bwd_row": lambda: N.gat_backward(*bargs, _phases=1, _buffers=bufs),
                         "bwd_col": lambda: N.gat_backward(*bargs, _phases=2, _buffers=bufs)}
            for kname, call in calls.items():
                call()
                pairs = []
                for _ in range(args.steps):
                    flush.fill_(1.0)
                    a, b_ = ev(), ev()
                    a.record()
                    call()
                    b_.record()
                    pairs.append((a, b_))
                torch.cuda.synchronize()
                kern_ms[kname] = [a.elapsed_time(b_) for a, b_ in pairs]
            for r, t in zip(recs, kern_ms["fwd"]):
                r["fwd_ms"] = t
        step_ms = [r["s"].elapsed_time(r["e"]) for r in recs]
        if kern_ms["fwd"]:
            fwd_ms = kern_ms["fwd"]
        elif graph is None:
            fwd_ms = [r["f0"].elapsed_time(r["f1"]) for r in recs]
        else:
            fwd_ms = [0.0] * len(recs)  # light pass under graph replay: no per-kernel events
        timed_coll = world > 1 and graph is None  # collectives only exist in the eager (row partition) step
        ag_ms = [r["ag0"].elapsed_time(r["ag1"]) for r in recs] if timed_coll else [0.0] * len(recs)
        rs_ms = [r["rs0"].elapsed_time(r["rs1"]) for r in recs] if timed_coll else [0.0] * len(recs)
        ms_local = sum(step_ms) / len(step_ms)

        # ---- e2e: public autograd API with host operands ------------------------------
        h_out = {}

        up, down = torch.cuda.Stream(), torch.cuda.Stream()

        def step_e2e():
            """Host operands in, host results out, through the public autograd operators.  The
            upstream gradient is uploaded on a second stream while the forward runs and the forward
            output is downloaded on a third while the backward runs (PCIe is full duplex)."""
            cur = torch.cuda.current_stream()
            dd = {k: v.to(dev, non_blocking=True) for k, v in h_in.items() if k != "dO"}
            up.wait_stream(cur)
            with torch.cuda.stream(up):
                dd["dO"] = h_in["dO"].to(dev, non_blocking=True)
            if conv == "gt":
                K, V = halo.gather_pair(dd["K"], dd["V"], None)
                Q = dd["Q"].requires_grad_()
                K.requires_grad_()
                V.requires_grad_()
                out = GTConvFuse_hyper(rows, row_ptr, col_ind, val, col_ptr, row_ind, val_idx, smem, Q, K, V)
                _download(out.detach(), "out", cur)
                cur.wait_stream(up)
                out.backward(dd["dO"])
                gk, gv = halo.reduce_pair(K.grad, V.grad, None)
                res = {"out": out.detach(), "gQ": Q.grad, "gK": gk, "gV": gv}
            else:
                F, ac = halo.gather_pair(dd["F"], dd["ac"], None)
                ar = dd["ar"].requires_grad_()
                ac.requires_grad_()
                F.requires_grad_()
                out = GATConvFuse(ar, ac, row_ptr, col_ind, col_ptr, row_ind, val_idx, 0.2, F, 0.0)
                _download(out.detach(), "out", cur)
                cur.wait_stream(up)
                out.backward(dd["dO"])
                gf, gc = halo.reduce_pair(F.grad, ac.grad, None)
                res = {"out": out.detach(), "gF": gf, "g_ar": ar.grad, "g_ac": gc}
            for k, v in res.items():
                if k == "out":
                    continue  # already on its way (download stream)
                if k not in h_out:
                    h_out[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                h_out[k].copy_(v, non_blocking=True)
            cur.wait_stream(down)
            dd["dO"].record_stream(cur)
            return res

        def _download(t, key, cur):
            if key not in h_out:
                h_out[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
            down.wait_stream(cur)
            with torch.cuda.stream(down):
                h_out[key].copy_(t, non_blocking=True)
            t.record_stream(down)

        for _ in range(1 if light else 3):
            step_e2e()
        barrier()
        e2e_ms = []
        for _ in range(1 if light else args.steps):
            flush.fill_(1.0)
            s, e = ev(), ev()
            s.record()
            step_e2e()
            e.record()
            e.synchronize()
            e2e_ms.append(s.elapsed_time(e))
        barrier()
        clock_info = clocks.stop() if clocks else None
        e2e_local = sum(e2e_ms) / len(e2e_ms)
        h2d = sum(v.numel() * v.element_size() for v in h_in.values())
        d2h = sum(v.numel() * v.element_size() for v in h_out.values())

        # ---- max over ranks -----------------------------------------------------------
        stats = torch.tensor([ms_local, e2e_local, sum(fwd_ms) / len(fwd_ms), sum(ag_ms) / len(ag_ms),
                              sum(rs_ms) / len(rs_ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        ms, e2e_t, fwd_t, ag_t, rs_t = (float(x) for x in stats.cpu())
        fwd_t = max(fwd_t, 1e-9)
        tot = torch.tensor([float(e_local), float(n_rows)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        e_all, n_all = (float(x) for x in tot.cpu())
        units = e_all * dim  # edges*dim processed by all ranks per step

        line = None
        if rank == 0:
            peak, peak_src = hbm_peak()
            fwd_bytes = alg_bytes(conv, "fwd", n_rows, e_local, dim)
            step_bytes = alg_bytes(conv, "fwd+bwd", n_rows, e_local, dim)
            staged = e_local <= 16 * max(n_rows, 1)  # abi_common.h: want_staged (mean degree <= 16)
            knames = ({"fwd": "gat_fwd_staged_kernel", "bwd_row": "gat_bwd_row_staged_kernel",
                       "bwd_col": "gat_bwd_col_staged_kernel"} if staged else
                      {"fwd": "gat_fwd_kernel", "bwd_row": "gat_bwd_row_kernel", "bwd_col": "gat_bwd_col_kernel"}) \
                if conv == "gat" else {"fwd": "dot_fwd_kernel", "bwd_row": "gt_bwd_row_kernel",
                                       "bwd_col": "gt_bwd_col_kernel"}
            fwd_kernel = knames["fwd"]
            kbytes = kernel_alg_bytes(conv, n_rows, e_local, dim)
            traffic = {}
            try:
                with open(os.path.join(ROOT, "profiles", "r01i_traffic.json")) as fh:
                    traffic = json.load(fh).get(name, {})
            except Exception:
                pass
            kernels = {}
            for k, ts in kern_ms.items():
                if ts:
                    t = sum(ts) / len(ts)
                    kernels[knames[k]] = {"ms": t, "algorithmic_bytes": kbytes[k],
                                          "achieved": kbytes[k] / (t * 1e-3) / 1e9,
                                          "frac": kbytes[k] / (t * 1e-3) / 1e9 / peak,
                                          "traffic": (traffic.get(knames[k]) or {}).get("dram_bytes_per_launch")}
            dominant = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
            line = {
                "metric": METRIC, "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak" if (weak or world == 1) else "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": name, "conv": conv, "dim": dim, "heads": 1, "format": fmt,
                           "nodes": int(n_all), "edges": int(e_all), "baseline_config_index": cfg,
                           "partition": part.describe, "l2": "flushed between timed steps (256 MB fill)",
                           "launch": "cuda graph replay" if graph is not None else "eager (collectives in the step)",
                           "graph_sha256": g_full.sha256()[:16]},
                "e2e": {"value": units / (e2e_t * 1e-3), "unit": UNIT, "ms_per_step": e2e_t,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "dfgnn_b200.operators.%s (autograd Function) with pinned host operands; "
                               "index formats resident" % ("GTConvFuse_hyper" if conv == "gt" else "GATConvFuse")},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm",
                             "kernel": dominant or fwd_kernel,
                             "achieved": kernels[dominant]["achieved"] if dominant else fwd_bytes / (fwd_t * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s",
                             "frac": kernels[dominant]["frac"] if dominant else fwd_bytes / (fwd_t * 1e-3) / 1e9 / peak,
                             "traffic": kernels[dominant]["traffic"] if dominant else None,
                             "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                                               "profiles/r01i_traffic.json" if dominant and kernels[dominant]["traffic"] else None,
                             "peak_source": peak_src,
                             "kernel_ms": kernels[dominant]["ms"] if dominant else fwd_t,
                             "algorithmic_bytes": kernels[dominant]["algorithmic_bytes"] if dominant else fwd_bytes,
                             "kernels": kernels,
                             "step": {"algorithmic_bytes": step_bytes,
                                      "achieved": step_bytes / (ms * 1e-3) / 1e9,
                                      "frac": step_bytes / (ms * 1e-3) / 1e9 / peak}},
                "format_construction_ms": fmt_ms,
                "clocks": clock_info,
                "collectives": {"allgather_ms": ag_t, "reduce_scatter_ms": rs_t} if (world > 1 and halo.active) else None,
            }

        # ---- GPU reference (the reference's own kernels, sm_100a) + CPU baseline, N = 1 ---
        if rank == 0 and world == 1 and not args.no_ref and not light:
            line["gpu_reference"] = time_gpu_reference(name, conv, dim, dict(
                row_ptr=row_ptr, col_ind=col_ind, rows=rows, val=val, col_ptr=col_ptr, row_ind=row_ind,
                val_idx=val_idx), d_in, flush, args.steps, e_total)
        if rank == 0 and world == 1 and not args.no_cpu and not light:
            line["cpu_baseline"] = cpu_baseline(name, g_full)
        return line

    name = args.workload
    full_graph = "batch" not in WORKLOADS[name][3]
    if world > 1 and full_graph and name != "reddit-gt" and args.scaling == "auto":
        # A graph of this size is one GPU's worth of work: at N > 1 the headline
        # is weak scaling -- one such graph per GPU, no collective -- and the row-partitioned
        # (halo all-gather + reduce-scatter) numbers of ONE graph over N ranks ride along.
        strong = measure("strong", True)
        line = measure("weak", False)
        if rank == 0:
            line["row_partition"] = {k: strong[k] for k in ("ms_per_step", "value", "collectives", "scaling")}
            line["row_partition"]["partition"] = strong["config"]["partition"]
    else:
        line = measure(args.scaling, False)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(name, g_full):
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    if name == "reddit-gt":
        from dfgnn_b200 import graphs
        g = graphs.reddit_like(0.05)
        sample = f"reddit-shaped at scale 0.05 (N={g.num_nodes()}, E={g.num_edges()}), 2 steps"
    else:
        g = g_full
        sample = f"full workload (N={g.num_nodes()}, E={g.num_edges()}), 3 steps after 1 warm-up"
    val, dt, *_ = time_cpu(name, g, 2 if name == "reddit-gt" else 3, 1)
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "ms_per_step": dt * 1e3}


def time_gpu_reference(name, conv, dim, idx, d_in, flush, steps, e_total):
    """The reference's DFGNN CUDA kernels (unchanged, sm_100a) on the same resident inputs."""
    import torch
    from oracle import ref_gpu
    if not ref_gpu.available():
        return {"unavailable": "oracle/_ref not built"}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    res = {}
    try:
        if conv == "gt":
            ref = ref_gpu.fused_gtconv()
            hs = ref_gpu.hyper_smem(idx["row_ptr"])
            args = (idx["row_ptr"], idx["col_ind"], idx["rows"], idx["val"], idx["col_ptr"],
                    idx["row_ind"], idx["val_idx"])
            Q, K, V, dO = d_in["Q"], d_in["K"], d_in["V"], d_in["dO"]
            if hs <= 12288:
                def fwd_bwd():
                    out, attn = ref.gt_hyper_forward(*args, hs, Q, K, V)
                    return ref.gt_backward(*args, hs, Q, K, V, attn, dO)
                variants = {"fwd+bwd (gt_hyper_forward + gt_backward)": fwd_bwd,
                            "fwd (gt_hyper_inference)": lambda: ref.gt_hyper_inference(
                                idx["row_ptr"], idx["col_ind"], idx["rows"], idx["val"], hs, Q, K, V)}
            else:
                variants = {}
                res["note"] = ("hyper/backward kernels need %d floats of smem per 8-row block (> 48 KB): "
                               "outside the reference's envelope; tiling forward only" % hs)
            variants["fwd (gt_tiling_inference)"] = lambda: ref.gt_tiling_inference(
                idx["row_ptr"], idx["col_ind"], idx["val"], 128, Q, K, V)
        else:
            ref = ref_gpu.fused_gatconv()
            ar, ac, F, dO = d_in["ar"], d_in["ac"], d_in["F"], d_in["dO"]
            ss = ref_gpu.softmax_smem(idx["row_ptr"])
            rows = torch.repeat_interleave(
                torch.arange(idx["row_ptr"].numel() - 1, device=F.device, dtype=torch.int32),
                (idx["row_ptr"][1:] - idx["row_ptr"][:-1]).long())

            def fwd_bwd():
                out, emax, esum, emask = ref.gat_forward(ar, ac, idx["row_ptr"], idx["col_ind"], 0.2, F, 0.0)
                return ref.gat_backward(0.2, 0.0, idx["row_ptr"], idx["col_ind"], idx["col_ptr"],
                                        idx["row_ind"], idx["val_idx"], emax, esum, emask, F, ar, ac, dO)
            variants = {"fwd+bwd (gat_forward + gat_backward)": fwd_bwd,
                        "fwd (gat_inference_softmax)": lambda: ref.gat_inference_softmax(
                            ss, ar, ac, idx["row_ptr"], idx["col_ind"], rows, 0.2, F)}
        for label, fn in variants.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(max(3, min(steps, 10))):
                flush.fill_(1.0)
                s, e = ev(), ev()
                s.record()
                fn()
                e.record()
                e.synchronize()
                ts.append(s.elapsed_time(e))
            ms = sorted(ts)[len(ts) // 2]  # median: gat_forward creates a cuRAND generator per call
            res[label] = {"ms": ms, "ms_min": min(ts), "edges_x_dim_per_s": e_total * dim / (ms * 1e-3)}
    except Exception as exc:  # the reference kernels abort on launch errors
        res["error"] = repr(exc)[:300]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="arxiv-gat", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ref", action="store_true", help="skip the reference-CUDA-kernel timing leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="batched workloads at N>1: weak (own batch per GPU, default) or strong "
                         "(one global batch sharded by graph); full graphs are always row-partitioned")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
