#!/usr/bin/env python3
"""Device time of the wiggle move sets, lookahead 1 and 2.  usage: wiggle_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine

W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]


def main():
    eng = Engine(0)
    eng.set_features(W6)
    rng = np.random.default_rng(3)
    print("%-36s %10s %12s %14s" % ("workload", "kernel_ms", "placements", "placements/s"))
    for la, n, T in ((1, 2048, 300), (2, 2048, 60)):
        w = np.array(W6V)[None, :] + rng.normal(0, 0.3, size=(n, 6))
        for ms, nm in ((4, "wiggle_top"), (8, "wiggle_drop")):
            best = None
            for _ in range(3):
                out = eng.play_games(w, 1, T=T, seed=5, seq_of_game=np.arange(n, dtype=np.int32), lookahead=la, move_set=ms)
                t = eng.last_timing()["kernel_ms"]
                best = t if best is None else min(best, t)
            pl = int(out["results"]["placements"].sum())
            print("%-36s %10.3f %12d %14.4e" % ("w6 %s L%d %d games T=%d" % (nm, la, n, T), best, pl, pl / (best * 1e-3)))


if __name__ == "__main__":
    main()
