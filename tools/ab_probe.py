"""A/B experiment helper: time the config-2 kernel and the latency probe with an alternative build of the library.
usage: ab_probe.py [path/to/alt.so]"""
import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.join(ROOT, sys.argv[1])
from tetris_ai_b200 import Engine
sys.path.insert(0, os.path.join(ROOT, "tools"))
from perf_probe import ce_populations, W6_NAMES, W6_VALS
eng = Engine(0); eng.set_features(W6_NAMES)
pops, seqs = ce_populations()
for lanes in (32, 64):
    best = 1e9
    for _ in range(7):
        out = eng.play_games(pops[1], 10, pieces=seqs, lanes=lanes)
        best = min(best, eng.last_timing()["kernel_ms"])
    print("ce_gen1 lanes %d: %.4f ms" % (lanes, best))
T = 2000
r = random.Random(0)
seq = np.array([[r.randrange(7) for _ in range(T)]], dtype=np.uint8)
for n in (1, 1000):
    w = np.tile(np.array(W6_VALS), (n, 1))
    best = 1e9
    for _ in range(4):
        eng.play_games(w, 1, pieces=seq, lanes=32)
        best = min(best, eng.last_timing()["kernel_ms"])
    print("identical games n=%d lanes 32: %.1f cycles/piece" % (n, best * 1e-3 * 1.965e9 / T))
n = 262144
rng = np.random.default_rng(1)
w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(n, 6))
best = 1e9
for _ in range(2):
    out = eng.play_games(w, 1, T=400, seed=5, seq_of_game=np.arange(n, dtype=np.int32))
    best = min(best, eng.last_timing()["kernel_ms"])
print("sweep 262144: %.3e placements/s (lanes %d)" % (int(out["results"]["placements"].sum()) / (best * 1e-3), eng.last_timing()["lanes"]))
