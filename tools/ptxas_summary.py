#!/usr/bin/env python
"""Summarise `-Xptxas -v` logs (taco_b200/lib/obj/*.ptxas.log): registers / shared memory / spills per kernel."""
import glob, os, re, subprocess, sys

here = os.path.dirname(os.path.abspath(__file__))
logs = sorted(glob.glob(os.path.join(here, "..", "taco_b200", "lib", "obj", "*.ptxas.log")))
rows = []
for log in logs:
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", txt):
        name, stack, sst, sld, regs, rest = m.groups()
        smem = re.search(r"(\d+) bytes smem", rest)
        rows.append((os.path.basename(log).split(".")[0], name, int(regs), int(smem.group(1)) if smem else 0, int(stack), int(sst)))
names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
flt = sys.argv[1] if len(sys.argv) > 1 else ""
print(f"{'file':10s} {'regs':>4s} {'smem':>6s} {'stack':>5s} {'spill':>5s}  kernel")
for r, n in zip(rows, names):
    n = re.sub(r"\(.*", "", n).replace("void tb::", "")
    if flt in n:
        print(f"{r[0]:10s} {r[2]:4d} {r[3]:6d} {r[4]:5d} {r[5]:5d}  {n}")
