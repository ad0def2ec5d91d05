#!/usr/bin/env python3
"""One play of N games on independent hashed streams (BASELINE configs[4] shape) for profiling.  usage: bin_one.py [N] [T]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
eng = Engine(0); eng.set_features(W6)
rng = np.random.default_rng(1)
w = np.array(W6V)[None, :] + rng.normal(0, 3.0, size=(n, 6))
out = eng.play_games(w, 1, T=T, seed=5, seq_of_game=np.arange(n, dtype=np.int32))
tm = eng.last_timing()
print("games %d T %d lanes %d launches %d kernel_ms %.3f placements %d pieces %d" % (n, T, tm["lanes"], tm["kernels_launched"], tm["kernel_ms"],
      int(out["results"]["placements"].sum()), int(out["results"]["pieces"].sum())))
