#!/usr/bin/env python3
"""One play of the BASELINE configs[2] population (C x 30 games on 30 shared sequences) for profiling.  usage: lock_one.py [C]"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine


def seeded_sequence(seed, T):
    """The canonical injected piece sequence: ONE random.Random(seed), T draws of randrange(7)."""
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
C = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G, T = 30, 1000
rng = np.random.default_rng(1)
w = np.array(W6V)[None, :] + rng.normal(0, 1.0, size=(C, 6))
seqs = np.array([seeded_sequence(100 + g, T) for g in range(G)], dtype=np.uint8)
eng = Engine(0); eng.set_features(W6)
r = eng.play_games(w, G, pieces=seqs)
tm = eng.last_timing()
print("C=%d kernel_ms %.3f lanes %d grid %d placements %d pieces %d" % (C, tm["kernel_ms"], tm["lanes"], tm["grid"], int(r["results"]["placements"].sum()), int(r["results"]["pieces"].sum())))
