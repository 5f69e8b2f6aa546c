"""Per-instruction view of one kernel of an .ncu-rep: executed count, stall samples, top stall.

    python tools/ncu_source.py REPORT KERNEL_REGEX [--min-exec N] [--view sass|cuda]
"""
import csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
view = "sass"
if "--view" in sys.argv:
    view = sys.argv[sys.argv.index("--view") + 1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"]
if view == "cuda":
    cmd += ["--print-source", "cuda"]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
blocks, cur = [], []
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur:
            blocks.append(cur)
        cur = [line]
    elif cur:
        cur.append(line)
if cur:
    blocks.append(cur)
b = blocks[0]
print(b[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
H = rows[0]
iS, iE, iN = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
tot_e = sum(int(r[iE] or 0) for r in rows[1:])
tot_s = sum(int(r[iN] or 0) for r in rows[1:])
print(f"total executed {tot_e}, samples {tot_s}")
for r in rows[1:]:
    e, n = int(r[iE] or 0), int(r[iN] or 0)
    st = sorted(((int(r[i] or 0), H[i]) for i in stall_cols), reverse=True)[:2]
    s = " ".join(f"{h[6:]}={v}" for v, h in st if v > 0)
    print(f"{e:9d} {100.0 * e / max(tot_e, 1):5.2f}% {n:6d} {100.0 * n / max(tot_s, 1):5.2f}%  {r[iS].strip()[:90]:90s} {s}")
