"""One throughput-regime launch (BASELINE configs[4] shape): N concurrent games, each its own candidate and hashed stream."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 400
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = Engine(0); eng.set_features(W6_NAMES)
rng = np.random.default_rng(1)
w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(n, 6))
best = 1e9
for _ in range(reps):
    out = eng.play_games(w, 1, T=T, seed=5, seq_of_game=np.arange(n, dtype=np.int32), lanes=lanes)
    best = min(best, eng.last_timing()["kernel_ms"])
pl = int(out["results"]["placements"].sum())
print("games %d T %d lanes %d kernel_ms %.3f placements %d pieces %d placements/s %.4e" % (n, T, eng.last_timing()["lanes"], best, pl, int(out["results"]["pieces"].sum()), pl / (best * 1e-3)))
