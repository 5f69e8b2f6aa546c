"""Human-readable digest of bench.py JSON lines:  python tools/show_bench.py FILE [FILE ...]"""
import json
import sys


def show(x):
    r = x["roofline"]
    name = x.get("workload") or x["config"]["workload"]
    cfg = x.get("config", x)
    print(f"{name:11s} N={x.get('n_gpus', '')} step {x['ms_per_step']:.4f} ms (eager {x['eager_ms_per_step']:.4f})  "
          f"value {x['value']:.3e}  [{cfg.get('launch')}] {x['scaling']}")
    for k, v in r["kernels"].items():
        fd = v.get("frac_dram")
        print(f"    {k:28s} {v['ms']:.4f} ms  frac {v['frac']:.2f}  compulsory {v['frac_compulsory']:.2f}  "
              f"dram {('%.2f' % fd) if fd else '-'}")
    s = r["step"]
    print(f"    step frac {s['frac']:.2f} compulsory {s['frac_compulsory']:.2f} dram {s.get('frac_dram')}")
    c = x.get("collectives")
    if c:
        print(f"    collectives: ag {c['allgather_ms']:.3f} rs {c['reduce_scatter_ms']:.3f} ms "
              f"share {c['share_of_step']:.2f} exposed {c['exposed_ms']:.3f} ms")
    g = x.get("gpu_reference") or {}
    for k, v in g.items():
        if isinstance(v, dict) and "ms" in v:
            print(f"    ref {k}: {v['ms']:.3f} ms")
    e = x.get("e2e")
    if e and e.get("ms_per_step"):
        print(f"    e2e {e['ms_per_step']:.3f} ms  host bytes / step time {e['host_bytes_over_step_gbs_per_gpu']:.1f} GB/s per GPU")


for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as ex:
        print(f, "ERR", ex)
        continue
    print("==", f)
    show(d)
    for w in d.get("workloads", []):
        if "error" in w:
            print(w)
        else:
            w = dict(w, n_gpus=d["n_gpus"])
            show(w)
