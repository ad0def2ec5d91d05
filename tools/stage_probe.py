#!/usr/bin/env python3
"""Kernel time of the BASELINE config-2 generation under different hand-over thresholds of the re-spreading stages
(TB200_STAGE_CUTS=a,b; TB200_NO_STAGES=1 = one launch).  usage: stage_probe.py  (spawns one process per setting)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import bench
from tetris_ai_b200 import Engine
pop, seqs = bench.canonical_population(100)
eng = Engine(0); eng.set_features(bench.W6_NAMES)
ts = []
for _ in range(12):
    r = eng.play_games(pop, 10, pieces=seqs)
    ts.append(eng.last_timing()["kernel_ms"])
ts = sorted(ts[2:])
print("%%.4f %%.4f %%d" %% (ts[0], ts[len(ts) // 2], eng.last_timing()["kernels_launched"]))
''' % ROOT


def main():
    settings = [("one launch", {"TB200_NO_STAGES": "1"})]
    for cuts in sys.argv[1:] or ["592,296", "592,0", "592,200", "592,250", "500,296", "450,250", "592,148"]:
        settings.append(("cuts " + cuts, {"TB200_STAGE_CUTS": cuts}))
    print("%-16s %10s %10s %8s" % ("setting", "min_ms", "median_ms", "launches"))
    for name, env in settings:
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, check=True).stdout.split()
        print("%-16s %10s %10s %8s" % (name, out[0], out[1], out[2]))


if __name__ == "__main__":
    main()
