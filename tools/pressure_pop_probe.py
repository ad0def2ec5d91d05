#!/usr/bin/env python3
"""move_with_pressure(move_drop) (simulator.py:88-120) on a CE population: the per-move time predicate inside the lockstep kernels
(default) against the per-game extended generic kernel (lanes = 8) and against the same population without the filter.
usage: pressure_pop_probe.py [C [G [T]]]"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
C = int(sys.argv[1]) if len(sys.argv) > 1 else 10000; G = int(sys.argv[2]) if len(sys.argv) > 2 else 30; T = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
def seeded_sequence(seed, T):
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]
rng = np.random.default_rng(1)
w = np.array(W6V)[None, :] + rng.normal(0, 1.0, size=(C, 6))
seqs = np.array([seeded_sequence(100 + g, T) for g in range(G)], dtype=np.uint8)
eng = Engine(0); eng.set_features(W6)
for name, kw in (("hard drops, lockstep (default)", {}), ("pressure (default parameters), lockstep (default)", {"pressure": {}}),
                 ("pressure, per-game generic kernel (lanes = 8)", {"pressure": {}, "lanes": 8})):
    best = 1e9
    for _ in range(2):
        out = eng.play_games(w, G, pieces=seqs, **kw)
        best = min(best, eng.last_timing()["kernel_ms"])
    pl = int(out["results"]["placements"].sum())
    print("%-52s C %d G %d T %d kernel_ms %.3f placements %d placements/s %.4e lanes %d launches %d" %
          (name, C, G, T, best, pl, pl / best / 1e-3, eng.last_timing()["lanes"], eng.last_timing()["kernels_launched"]), flush=True)
