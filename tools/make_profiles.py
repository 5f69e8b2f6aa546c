#!/usr/bin/env python
"""Turn the ncu outputs that gpurun brought back into the committed summaries under profiles/.

    python tools/make_profiles.py TAG launches.csv full_arxiv.ncu-rep [full_pattern.ncu-rep]

Writes profiles/TAG_launches_arxiv-gat.{csv,md}, profiles/TAG_ncu_full_*.txt and
profiles/TAG_traffic.json (per-launch DRAM bytes; bench.py copies them into roofline.traffic)."""
import collections, csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
prof = os.path.join(ROOT, "profiles")
shutil.copy(launches, os.path.join(prof, f"{tag}_launches_arxiv-gat.csv"))
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
H = rows[0]
ik, iv, iu = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += v
ours = {k for k in agg if "dfgnn::gat" in k or "dfgnn::dot" in k or "dfgnn::gt_" in k}
tot = sum(a[1] for a in agg.values())
tot_conv = sum(agg[k][1] for k in ours)
out = [f"# {tag}: ncu launch list, default bench (arxiv-gat)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-ref`",
       "(cold-cache, serialised per-launch times: compare SHARES; the timed step is a CUDA-graph replay of the same launches)", "",
       "| launches | avg us | share of all launches | share of the conv kernels | kernel |", "|---|---|---|---|---|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    sc = f"{100 * t / tot_conv:.1f}%" if k in ours else "-"
    out.append(f"| {n} | {t / n:.1f} | {100 * t / tot:.1f}% | {sc} | `{k[:100]}` |")
open(os.path.join(prof, f"{tag}_launches_arxiv-gat.md"), "w").write("\n".join(out) + "\n")


def traffic(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    Hh, U = rr[0], rr[1]
    res = {}
    for r in rr[2:]:
        name = r[Hh.index("Kernel Name")].split("(")[0].split("<")[0].replace("void ", "").replace("dfgnn::", "")

        def val(m):
            i = Hh.index(m)
            return float(r[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[U[i]]
        i = Hh.index("gpu__time_duration.sum")
        t = float(r[i].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}[U[i]]
        res.setdefault(name, []).append((val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), t))
    return {k: {"dram_bytes_per_launch": sum(x[0] for x in v) / len(v), "ncu_us": sum(x[1] for x in v) / len(v)}
            for k, v in res.items()}


T = {"source": f"ncu --set full --clock-control none, bench.py --steps 3 --warmup 3 --no-cpu --no-ref (profiles/{tag}_ncu_full_*.txt)"}
for rep in reps:
    wl = "pattern-gt" if "pattern" in rep else "arxiv-gat"
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(prof, f"{tag}_ncu_full_{wl}.txt"), "w").write(txt)
    T[wl] = traffic(rep)
json.dump(T, open(os.path.join(prof, f"{tag}_traffic.json"), "w"), indent=1)
print(json.dumps(T, indent=1))
