import json, sys
for tag in sys.argv[2:]:
    w = tag
    try:
        d=json.loads(open(f"gpurun_out/{sys.argv[1]}_{w}.json").read().strip().splitlines()[-1])
    except Exception as ex:
        print(w, "ERR", ex); print(open(f"gpurun_out/{sys.argv[1]}_{w}.err").read()[-1200:]); continue
    r=d["roofline"]
    ref=d.get("gpu_reference") or {}
    refs="; ".join(f"{k.split(' (')[0]} {v['ms']:.3f}" for k,v in ref.items() if isinstance(v,dict) and 'ms' in v)
    print(f"{w:11s} step {d['ms_per_step']:.4f} ms (frac {r['step']['frac']:.2f})  {r['kernel'][:22]} {r['kernel_ms']:.4f} ms (frac {r['frac']:.2f})  e2e {d['e2e']['ms_per_step']:.2f} ms | ref: {refs}")
