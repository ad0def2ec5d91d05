#!/usr/bin/env python3
"""Tuning of the lockstep kernel K2s on the BASELINE configs[2] shape: placements/s by lanes per game and stage length
(TB200_LOCK_LANES / TB200_LOCK_K; one process per setting).  usage: lock_probe.py C spread"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, random, numpy as np
sys.path.insert(0, %r)
from tetris_ai_b200 import Engine


def seeded_sequence(seed, T):
    """The canonical injected piece sequence (SURVEY.md §8c): ONE random.Random(seed), T draws of randrange(7)."""
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
C, spread = int(sys.argv[1]), float(sys.argv[2]); G, T = 30, 1000
rng = np.random.default_rng(1)
w = np.array(W6V)[None, :] + rng.normal(0, spread, size=(C, 6))
seqs = np.array([seeded_sequence(100 + g, T) for g in range(G)], dtype=np.uint8)
eng = Engine(0); eng.set_features(W6)
best = 1e9
for _ in range(3):
    r = eng.play_games(w, G, pieces=seqs)
    best = min(best, eng.last_timing()["kernel_ms"])
print("%%.3f %%.4e %%d" %% (best, int(r["results"]["placements"].sum()) / best / 1e-3, eng.last_timing()["lanes"]))
''' % ROOT
C, spread = sys.argv[1], sys.argv[2]
print("C=%s spread=%s" % (C, spread))
for lanes in (8, 4, 2, 1):
    for K in (64,):
        e = dict(os.environ); e["TB200_LOCK_LANES"] = str(lanes); e["TB200_LOCK_K"] = str(K)
        out = subprocess.run([sys.executable, "-c", CHILD, C, spread], env=e, capture_output=True, text=True, timeout=300).stdout.split()
        print("lanes %d K %3d  %9s ms  %s placements/s" % (lanes, K, out[0], out[1]))
