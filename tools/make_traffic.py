#!/usr/bin/env python
"""Turn what tools/profile_r02.sh brought back in gpurun_out/ into the committed evidence:

    python tools/make_traffic.py [TAG]        (default TAG r03)

  profiles/TAG_traffic.json          per kernel and workload: DRAM bytes per launch
                                     (dram__bytes_read.sum + dram__bytes_write.sum) and the ncu
                                     duration, for OUR kernels and for the reference's kernels
                                     launched in the same capture; `source_sha` = hash of the kernel
                                     sources the capture was taken from (bench.py reports `traffic`
                                     only while it matches the running sources)
  profiles/TAG_ncu_full_<wl>.txt     --set full summary of our kernels (tools/ncu_summary.py)
  profiles/TAG_launches_arxiv-gat.*  launch list of the default bench command
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TAG = sys.argv[1] if len(sys.argv) > 1 else "r03"
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(kname):
    k = kname.replace("void ", "")
    k = k.split("(")[0].split("<")[0]
    return k.split("::")[-1]


def parse(path):
    """-> {kernel: [ {metric: value} per launch ]} in launch order."""
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    if not rows:
        return {}
    H = rows[0]
    iid, ik, im, iu, iv = (H.index(c) for c in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
        except ValueError:
            continue
        per.setdefault((r[iid], short(r[ik])), {})[r[im]] = v
    out = collections.OrderedDict()
    for (_, k), m in per.items():
        out.setdefault(k, []).append(m)
    return out


def main():
    from dfgnn_b200 import _lib
    T = {"source_sha": _lib.source_sha(), "tag": TAG,
         "how": "tools/profile_r02.sh: ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
                "dram__bytes_write.sum --clock-control none over `python bench.py --workload W --profile "
                "--profile-ref` (1 warm-up + 2 eager steps of ours, 4 calls of each reference variant); "
                "per-launch averages over the launches AFTER the warm-up step; launches are cold-cache and "
                "serialised under ncu: compare shares and bytes, not absolute times",
         "workloads": {}, "reference_kernels": {}}
    try:
        T["commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True,
                                     text=True).stdout.strip()
    except Exception:
        pass
    for wl in ("arxiv-gat", "pattern-gt", "voc-gt", "reddit-gt"):
        path = os.path.join(OUT, f"{TAG}_traffic_{wl}.csv")
        if not os.path.exists(path):
            continue
        per = parse(path)
        ours, ref = {}, {}
        for k, launches in per.items():
            is_ours = k.startswith(("gat_fwd", "gat_bwd", "dot_fwd", "gt_bwd", "gt_block", "gt_dense", "block_attn_dense"))
            use = launches[len(launches) // 3:] if (is_ours and len(launches) >= 3) else launches
            # big-tile fallback launches of the staged GAT path return at once: keep them apart
            rec = {"launches": len(launches),
                   "dram_bytes_per_launch": sum(m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
                                                for m in use) / len(use),
                   "ncu_us": sum(m.get("gpu__time_duration.sum", 0) for m in use) / len(use)}
            (ours if is_ours else ref)[k] = rec
        T["workloads"][wl] = ours
        T["reference_kernels"][wl] = ref
        rep = os.path.join(OUT, f"{TAG}_full_{wl}.ncu-rep")
        if os.path.exists(rep):
            txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep],
                                 capture_output=True, text=True).stdout
            # keep the LAST captured instance of every kernel (the timed steps, not the warm-up)
            blocks = txt.split("----- ")
            last = collections.OrderedDict()
            for b in blocks[1:]:
                last[b.split("\n", 1)[0]] = b
            open(os.path.join(PROF, f"{TAG}_ncu_full_{wl}.txt"), "w").write(
                "".join("----- " + b for b in last.values()))
    json.dump(T, open(os.path.join(PROF, f"{TAG}_traffic.json"), "w"), indent=1)
    lst = os.path.join(OUT, f"{TAG}_launches_arxiv-gat.csv")
    if os.path.exists(lst):
        per = parse(lst)
        tot = sum(m.get("gpu__time_duration.sum", 0) for ls in per.values() for m in ls)
        conv = {k: ls for k, ls in per.items() if k.startswith(("gat_", "dot_fwd", "gt_bwd", "gt_block", "gt_dense", "block_attn"))}
        tot_conv = sum(m.get("gpu__time_duration.sum", 0) for ls in conv.values() for m in ls)
        lines = [f"# {TAG}: ncu launch list, default bench (arxiv-gat)", "",
                 "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 3 "
                 "--warmup 3 --no-cpu --no-ref --no-extras`",
                 "(cold-cache, serialised per-launch times: compare SHARES; the timed step is a CUDA-graph "
                 "replay of the same launches)", "",
                 "| launches | avg us | share of all launches | share of the conv kernels | kernel |",
                 "|---|---|---|---|---|"]
        for k, ls in sorted(per.items(), key=lambda kv: -sum(m.get("gpu__time_duration.sum", 0) for m in kv[1])):
            t = sum(m.get("gpu__time_duration.sum", 0) for m in ls)
            sc = f"{100 * t / tot_conv:.1f}%" if k in conv else "-"
            lines.append(f"| {len(ls)} | {t / len(ls):.1f} | {100 * t / tot:.1f}% | {sc} | `{k}` |")
        open(os.path.join(PROF, f"{TAG}_launches_arxiv-gat.md"), "w").write("\n".join(lines) + "\n")
        import shutil
        shutil.copy(lst, os.path.join(PROF, f"{TAG}_launches_arxiv-gat.csv"))
    print(json.dumps(T, indent=1)[:3000])


if __name__ == "__main__":
    main()
