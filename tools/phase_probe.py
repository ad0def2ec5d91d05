"""One-off experiment: per-phase cycle counts of the w32 kernel (needs a -DTB_PHASES build at gpurun_out/libphases.so)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "bin", "libphases.so")
from tetris_ai_b200 import Engine
sys.path.insert(0, os.path.join(ROOT, "tools"))
from perf_probe import ce_populations, W6_NAMES
eng = Engine(0); eng.set_features(W6_NAMES)
pops, seqs = ce_populations()
L = _lib.load()
out = (ctypes.c_ulonglong * 8)()
for gi in (1,):
    for rep in range(2):
        L.tb200_debug_phases(out, 1)
        r = eng.play_games(pops[gi], 10, pieces=seqs, lanes=32)
        L.tb200_debug_phases(out, 1)
    v = list(out); pieces = v[6]
    names = ["prologue(fetch,desc,placeA)", "eval single", "eval dual", "argmax", "commit", "loop-top"]
    tot = sum(v[:6])
    print("kernel_ms", eng.last_timing()["kernel_ms"], "pieces", pieces, "cycles/piece", tot / pieces)
    for n, x in zip(names, v[:6]):
        print("  %-30s %8.1f cycles/piece  %5.1f%%" % (n, x / pieces, 100 * x / tot))
