"""One-off experiment: per-phase cycle counts of the one-warp-per-game w6 kernel (needs a -DTB_PHASES build at bin/libphases.so)."""
import ctypes, os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "bin", "libphases.so")
from tetris_ai_b200 import Engine
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
eng = Engine(0); eng.set_features(W6_NAMES)
L = _lib.load()
out = (ctypes.c_ulonglong * 8)()
T = 2000
r = random.Random(0)
seq = np.array([[r.randrange(7) for _ in range(T)]], dtype=np.uint8)
for n in (1, 1000):
    w = np.tile(np.array(W6_VALS), (n, 1))
    for rep in range(2):
        L.tb200_debug_phases(out, 1)
        res = eng.play_games(w, 1, pieces=seq, lanes=32)
        L.tb200_debug_phases(out, 1)
    v = list(out); pieces = v[6]
    names = ["prologue(fetch,desc prefetch)", "eval single (pre+post)", "eval dual", "argmax", "commit", "loop-top"]
    tot = sum(v[:6])
    print("games", n, "kernel_ms", eng.last_timing()["kernel_ms"], "pieces", pieces, "cycles/piece", tot / pieces)
    for nm, x in zip(names, v[:6]):
        print("  %-30s %8.1f cycles/piece  %5.1f%%" % (nm, x / pieces, 100 * x / tot))
