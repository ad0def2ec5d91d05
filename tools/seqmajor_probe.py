#!/usr/bin/env python3
"""BASELINE configs[2] shape on one GPU shard: C candidates x 30 games on 30 shared piece sequences; placements/s by lanes per game.
TB200_NO_SEQ_MAJOR=1 restores game-index ticket order.  usage: seqmajor_probe.py [C] [spread]"""
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
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
spread = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
G, T = 30, 1000
rng = np.random.default_rng(1)
w = np.array(W6V)[None, :] + rng.normal(0, spread, size=(C, 6))
seqs = np.array([seeded_sequence(100 + g, T) for g in range(G)], dtype=np.uint8)
eng = Engine(0); eng.set_features(W6)
print("C=%d G=%d T=%d spread=%.1f seq_major=%s" % (C, G, T, spread, "off" if os.environ.get("TB200_NO_SEQ_MAJOR") else "on"))
ref = None
for lanes in (0, 16, 8, 4, 2, 1):
    best = 1e9
    for _ in range(2):
        r = eng.play_games(w, G, pieces=seqs, lanes=lanes)
        best = min(best, eng.last_timing()["kernel_ms"])
    res = r["results"]
    if ref is None: ref = res.copy()
    assert (res == ref).all()
    pl = int(res["placements"].sum())
    print("lanes %2d (used %2d)  %8.3f ms  %.4e placements/s   mean game length %.0f" % (lanes, eng.last_timing()["lanes"], best, pl / best / 1e-3, res["pieces"].mean()))
