#!/usr/bin/env python3
"""K2b (piece-binned kernel, independent streams) against K2g: parity of every result and kernel time.
usage: bin_probe.py [games] [T]   (TB200_BIN_LANES = 1 / 2 / 4 / 8 selects the lane width of K2b)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
eng = Engine(0); eng.set_features(W6)
rng = np.random.default_rng(1)
w = np.array(W6V)[None, :] + rng.normal(0, 3.0, size=(n, 6))
ids = np.arange(n, dtype=np.int32)
def run(lanes):
    best = 1e9
    for _ in range(2):
        out = eng.play_games(w, 1, T=T, seed=5, seq_of_game=ids, lanes=lanes)
        best = min(best, eng.last_timing()["kernel_ms"])
    return out["results"], best, eng.last_timing()
a, ta, tma = run(0)
b, tb, tmb = run(2)
pl = int(a["placements"].sum())
print("games %d T %d | auto: lanes %d launches %d %.3f ms %.4e placements/s | K2g<2>: %.3f ms %.4e | identical %s" % (
    n, T, tma["lanes"], tma["kernels_launched"], ta, pl / ta / 1e-3, tb, pl / tb / 1e-3, bool((a == b).all())))
