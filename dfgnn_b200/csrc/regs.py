"""Summarise `-Xptxas -v` logs written by the Makefile (registers / spills per kernel)."""
import glob, os, re, subprocess, sys

here = os.path.dirname(os.path.abspath(__file__))
for fn in sorted(glob.glob(os.path.join(here, "_build", "*.ptxas.log"))):
    txt = open(fn).read()
    ents = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, "
                      r"(\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt)
    if not ents:
        continue
    print(os.path.basename(fn))
    names = subprocess.run(["c++filt"] + [e[0] for e in ents], capture_output=True, text=True).stdout.split("\n")
    for (name, stack, ss, sl, regs), dem in zip(ents, names):
        dem = re.sub(r"dfgnn::|\(.*", "", dem)[:100]
        print(f"  {regs:>4} regs  spill {ss:>4}/{sl:<4} {dem}")
