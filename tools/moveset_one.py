"""One launch for ncu: w6 games with one of the SURVEY §8f move sets (1 = overhang slides, 4 = move_wiggle, 8 = move_with_wiggle(move_drop),
+2 = pressure filter).  usage: moveset_one.py <move_set> [games] [T]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
ms = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192; T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
eng = Engine(0); eng.set_features(W6)
rng = np.random.default_rng(3)
w = np.array(W6V)[None, :] + rng.normal(0, 1.0, size=(n, 6))
pressure = {"avg_lat": 250.9, "resp_time": 79.8, "move_eff": 1.23} if ms & 2 else None
best = 1e9
for _ in range(2):
    out = eng.play_games(w, 1, T=T, seed=5, seq_of_game=np.arange(n, dtype=np.int32), move_set=ms & ~2, pressure=pressure)
    best = min(best, eng.last_timing()["kernel_ms"])
pl = int(out["results"]["placements"].sum())
print("move_set %d games %d T %d kernel_ms %.3f placements %d placements/s %.4e lanes %d" % (ms, n, T, best, pl, pl / best / 1e-3, eng.last_timing()["lanes"]))
