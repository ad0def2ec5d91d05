"""Smallest case that touches every kernel, for compute-sanitizer (one tool per gpurun call)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
eng = Engine(0)
rng = np.random.default_rng(0)
eng.set_features(W6)
w = rng.normal(0, 5, size=(6, 6))
for lanes in (64, 32, 16, 8):
    r = eng.play_games(w, 2, T=60, seed=1, lanes=lanes, want_trace=True, want_scores=True, want_final=True)
r2 = eng.play_games(w, 2, T=20, seed=1, lookahead=2, lanes=8, want_trace=True)
eng.set_features(list(range(45)))
w45 = rng.normal(0, 1, size=(3, 45))
for lanes in (32, 8):
    eng.play_games(w45, 2, T=30, seed=2, lanes=lanes)
eng.play_games(w45, 1, T=12, seed=2, lookahead=2, lanes=16)
eng.set_features(["pits", "max_ht", "mean_pit_depth"])
eng.play_games(rng.normal(0, 1, size=(3, 3)), 2, T=30, seed=3, lanes=32)
boards = np.zeros((5, 20), dtype=np.uint16); boards[:, 19] = 0x3FE
eng.best_move(boards, [0, 1, 2, 3, 4], [1, 2, 3, 4, 5], rng.normal(0, 1, size=3), lookahead=2)
eng.eval_features(boards, [0, 1, 2, 3, 4], [1, 0, 0, 0, 0], [0, 1, 2, 3, 4])
eng.hashed_pieces(1, 0, 3, 50)
print("sanitize case done", int(r["results"]["pieces"].sum()))
