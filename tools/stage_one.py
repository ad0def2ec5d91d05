"""One play of the headline workload (BASELINE configs[1] generation 1: 100 x 10 games) for ncu: the re-spreading stages = three play launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tetris_ai_b200 import Engine
pop, seqs = bench.canonical_population(100)
eng = Engine(0); eng.set_features(bench.W6_NAMES)
r = eng.play_games(pop, 10, pieces=seqs)
print("kernel_ms %.4f launches %d placements %d pieces %d" % (eng.last_timing()["kernel_ms"], eng.last_timing()["kernels_launched"], int(r["results"]["placements"].sum()), int(r["results"]["pieces"].sum())))
