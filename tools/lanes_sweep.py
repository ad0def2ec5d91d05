"""Best lanes-per-game for each population size (w6, lookahead 1, hashed streams, T=400)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
eng = Engine(0); eng.set_features(W6_NAMES)
rng = np.random.default_rng(1)
print("%8s " % "games" + " ".join("%9s" % ("L%d" % l) for l in (64, 32, 16, 8, 4, 2, 1)))
for n in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 524288):
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(n, 6))
    row = []
    for lanes in (64, 32, 16, 8, 4, 2, 1):
        if (lanes >= 32 and n > 32768) or (lanes <= 2 and n < 4096):
            row.append("        -"); continue
        best = 1e9
        for _ in range(2):
            out = eng.play_games(w, 1, T=400, seed=5, seq_of_game=np.arange(n, dtype=np.int32), lanes=lanes)
            best = min(best, eng.last_timing()["kernel_ms"])
        row.append("%9.2e" % (int(out["results"]["placements"].sum()) / (best * 1e-3)))
    print("%8d " % n + " ".join(row), flush=True)
