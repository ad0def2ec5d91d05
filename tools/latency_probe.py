"""Per-piece latency probe: identical w6 games (all survive to T), kernel_ms / T for several population sizes."""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
eng = Engine(0); eng.set_features(W6_NAMES)
T = 2000
r = random.Random(0)
seq = np.array([[r.randrange(7) for _ in range(T)]], dtype=np.uint8)
print("%8s %6s %10s %12s %14s" % ("games", "lanes", "kernel_ms", "cyc/piece", "placements/s"))
for n in (592, 1000):
    w = np.tile(np.array(W6_VALS), (n, 1))
    for lanes in (32, 0):
        best = 1e9
        for _ in range(3):
            out = eng.play_games(w, 1, pieces=seq, lanes=lanes)
            best = min(best, eng.last_timing()["kernel_ms"])
        res = out["results"]
        assert int(res["pieces"].min()) == T
        print("%8d %6d %10.3f %12.1f %14.4e" % (n, lanes, best, best * 1e-3 * 1.965e9 / T, res["placements"].sum() / (best * 1e-3)))
