#!/usr/bin/env python3
"""SASS opcode histogram per kernel of libtetris_b200.so (cuobjdump -sass): the no-FMA-contraction proof (DFMA must be 0 wherever
the canonical fp64 scorer runs; it appears only where __ddiv_rn expands: the all-45 / run-time-list feature code and fitness_kernel)
and the POPC / FLO / fp64 counts per kernel.  usage: sass_opcodes.py [lib] > profiles/sass_opcodes_r02.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tetris_ai_b200", "libtetris_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
cols = ["DFMA", "DMUL", "DADD", "POPC", "FLO", "IDP", "SHFL", "REDUX", "CREDUX", "BAR", "LDS", "LDG"]
print("# %s: %d kernels; static SASS instruction counts (whole kernel)" % (os.path.basename(lib), len(hist)))
print("%-110s %7s " % ("kernel", "total") + " ".join("%6s" % c for c in cols))
for k, h in hist.items():
    name = demangle(k)
    name = re.sub(r"\(tb::PlayParams\)$", "", name)
    print("%-110s %7d " % (name[:110], sum(h.values())) + " ".join("%6d" % h[c] for c in cols))
w6 = [k for k in hist if re.search(r"w6fast|w6pair|w6quad|w6grp|w6bin|w6lockwin|w6lockauto", k) or re.search(r"w6lock_kernelILi\dELi1E", k)]
print("# DFMA in the w6 kernels (must be 0): %d over %d kernels" % (sum(hist[k]["DFMA"] for k in w6), len(w6)))
