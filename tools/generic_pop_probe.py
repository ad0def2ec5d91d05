#!/usr/bin/env python3
"""BASELINE configs[2] shape with feature lists other than w6: placements/s of the generic kernels.  usage: generic_pop_probe.py [C] [T]"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
from oracle import pyoracle


def seeded_sequence(seed, T):
    """The canonical injected piece sequence: ONE random.Random(seed), T draws of randrange(7)."""
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


C = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
G = 30
seqs = np.array([seeded_sequence(100 + g, T) for g in range(G)], dtype=np.uint8)
eng = Engine(0)
rng = np.random.default_rng(2)
sets = {
    "w6": (["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"], [-4.5, 3.42, -3.22, -9.35, -7.9, -3.39]),
    "8 features": (["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells", "max_ht", "jaggedness"], [-4.5, 3.42, -3.22, -9.35, -7.9, -3.39, -0.5, -0.3]),
    "all 45": (list(pyoracle.FEATURE_NAMES), None),
}
for name, (feats, base) in sets.items():
    eng.set_features(feats)
    if base is None:
        base = [0.0] * 45
        for n, v in zip(sets["w6"][0], sets["w6"][1]):
            base[feats.index(n)] = v
    w = np.array(base)[None, :] + rng.normal(0, 0.3, size=(C, len(feats)))
    best = 1e9
    for _ in range(2):
        r = eng.play_games(w, G, pieces=seqs)
        best = min(best, eng.last_timing()["kernel_ms"])
    pl = int(r["results"]["placements"].sum())
    print("%-12s C=%d T=%d lanes %2d  %9.3f ms  %.4e placements/s  mean length %.0f" % (name, C, T, eng.last_timing()["lanes"], best, pl / best / 1e-3, r["results"]["pieces"].mean()))
