#!/usr/bin/env python3
"""K2q (four warps per game, three lanes per placement) against K2f / K2p:
  (1) per-piece latency of all-surviving identical w6 games by games in flight and kernel (lanes 32 = K2f, 64 = K2p, 128 = K2q);
  (2) kernel time of the BASELINE config-2 generation (bench.py's headline step) under different stage schedules
      (TB200_QUAD=1 = K2q as the sparse stages; TB200_STAGE_CUTS=a,b,c = hand-over thresholds of stages 0 / 1 / 2).
usage: python tools/quad_probe.py  (spawns one process per setting for part 2)"""
import os, subprocess, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import bench
from tetris_ai_b200 import Engine
pop, seqs = bench.canonical_population(100)
eng = Engine(0); eng.set_features(bench.W6_NAMES)
ts = []
for _ in range(14):
    r = eng.play_games(pop, 10, pieces=seqs)
    ts.append(eng.last_timing()["kernel_ms"])
ts = sorted(ts[2:])
print("%%.4f %%.4f %%d %%d" %% (ts[0], ts[len(ts) // 2], eng.last_timing()["kernels_launched"], int(r["results"]["placements"].sum())))
''' % ROOT


def latency():
    from tetris_ai_b200 import Engine
    W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
    W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
    eng = Engine(0); eng.set_features(W6_NAMES)
    T = 2000
    r = random.Random(0)
    seq = np.array([[r.randrange(7) for _ in range(T)]], dtype=np.uint8)
    print("%8s %6s %10s %12s %14s" % ("games", "lanes", "kernel_ms", "cyc/piece", "placements/s"))
    for n in (74, 148, 296, 592):
        w = np.tile(np.array(W6_VALS), (n, 1))
        for lanes in (32, 64, 128):
            best = 1e9
            for _ in range(3):
                out = eng.play_games(w, 1, pieces=seq, lanes=lanes)
                best = min(best, eng.last_timing()["kernel_ms"])
            res = out["results"]
            assert int(res["pieces"].min()) == T
            print("%8d %6d %10.3f %12.1f %14.4e" % (n, lanes, best, best * 1e-3 * 1.965e9 / T, res["placements"].sum() / (best * 1e-3)), flush=True)
    eng.close()


def main():
    latency()
    settings = [("default (K2p last)", {}), ("K2q stages", {"TB200_QUAD": "1"})]
    for cuts in sys.argv[1:] or ["592,296,0", "592,296,100", "592,400,148", "592,444,148", "592,592,148", "592,592,296", "500,350,148"]:
        settings.append(("K2q cuts " + cuts, {"TB200_STAGE_CUTS": cuts, "TB200_QUAD": "1"}))
    print("%-20s %10s %10s %8s %12s" % ("setting", "min_ms", "median_ms", "launches", "placements"))
    for name, env in settings:
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, check=True).stdout.split()
        print("%-20s %10s %10s %8s %12s" % (name, out[0], out[1], out[2], out[3]), flush=True)


if __name__ == "__main__":
    main()
