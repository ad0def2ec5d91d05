"""One lockstep launch of a run-time specialised kernel for ncu: 8 features, C candidates x 30 games (T = 300).  usage: jit_one.py [C]"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
C = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
G, T = 30, 300
names = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells", "max_ht", "jaggedness"]
base = [-4.5, 3.42, -3.22, -9.35, -7.9, -3.39, -0.5, -0.3]
def seq(seed):
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]
seqs = np.array([seq(100 + g) for g in range(G)], dtype=np.uint8)
rng = np.random.default_rng(2)
w = np.array(base)[None, :] + rng.normal(0, 0.3, size=(C, len(names)))
eng = Engine(0); eng.set_features(names)
best = 1e9
for _ in range(2):
    r = eng.play_games(w, G, pieces=seqs)
    best = min(best, eng.last_timing()["kernel_ms"])
pl = int(r["results"]["placements"].sum())
print("C=%d kernel_ms %.3f placements %d placements/s %.4e lanes %d specialised %d" % (C, best, pl, pl / best / 1e-3, eng.last_timing()["lanes"], eng.last_timing()["specialised"]))
