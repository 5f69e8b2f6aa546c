"""SASS opcode histogram per kernel of the built library (evidence for profiles/):
    python tools/sass_histogram.py > profiles/rNN_sass_histogram.md
Lists, for every kernel of libdfgnn_b200.so, the instruction count, the opcodes that prove the
Blackwell paths (UTCxMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP =
cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async, HMMA = mma.sync) and the ten most frequent opcodes."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "dfgnn_b200", "libdfgnn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, hist = None, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCQMMA", "UTCHMMA", "UTCMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "HMMA", "MUFU", "ATOMS", "ATOMG", "RED", "SHFL", "BAR", "LDL", "STL")
print("# SASS opcode histogram of dfgnn_b200/libdfgnn_b200.so (sm_100a, `tools/sass_histogram.py`)\n")
print("| kernel | instructions | Blackwell / notable opcodes | ten most frequent |")
print("|---|---|---|---|")
for k, h in sorted(hist.items(), key=lambda kv: -sum(kv[1].values())):
    name = demangle(k)
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("dfgnn::", "")
    tot = sum(h.values())
    if tot < 200:
        continue
    seen, parts = set(), []
    for op in KEY:
        n = sum(v for o, v in h.items() if o.startswith(op))
        if n and op not in seen:
            parts.append(f"{op}* {n}")
            seen.add(op)
    top = ", ".join(f"{o} {v}" for o, v in h.most_common(10))
    print(f"| `{name[:110]}` | {tot} | {', '.join(parts)} | {top} |")
