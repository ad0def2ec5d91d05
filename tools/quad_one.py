"""One launch of one latency kernel for ncu: 148 identical all-surviving w6 games, T pieces.  usage: quad_one.py <lanes> [T]"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
lanes = int(sys.argv[1]); T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
eng = Engine(0); eng.set_features(W6_NAMES)
r = random.Random(0)
seq = np.array([[r.randrange(7) for _ in range(T)]], dtype=np.uint8)
w = np.tile(np.array(W6_VALS), (148, 1))
out = eng.play_games(w, 1, pieces=seq, lanes=lanes)
print(lanes, eng.last_timing()["kernel_ms"], int(out["results"]["pieces"].min()), int(out["results"]["placements"].sum()))
