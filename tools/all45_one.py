"""BASELINE configs[3] shape: two-piece lookahead over the full 45-feature set (weights = w6 + N(0, 0.3) on all 45)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
from tetris_ai_b200.tetris.features import FEATURE_ID
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
la = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0
eng = Engine(0); eng.set_features(list(range(45)))
rng = np.random.default_rng(1)
w = rng.normal(0, 0.3, size=(n, 45))
for i, nm in enumerate(W6_NAMES):
    w[:, FEATURE_ID[nm]] += W6_VALS[i]
best = 1e9
for _ in range(2):
    out = eng.play_games(w, 1, T=T, seed=5, seq_of_game=np.arange(n, dtype=np.int32), lookahead=la, lanes=lanes)
    best = min(best, eng.last_timing()["kernel_ms"])
pl = int(out["results"]["placements"].sum())
print("all45 L%d games %d T %d lanes %d kernel_ms %.3f placements %d pieces %d placements/s %.4e" % (la, n, T, eng.last_timing()["lanes"], best, pl, int(out["results"]["pieces"].sum()), pl / (best * 1e-3)))
