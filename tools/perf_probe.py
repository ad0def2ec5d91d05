#!/usr/bin/env python3
"""Quick on-box timing probe (not the bench): kernel time of K2 for several workloads / lane widths."""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine  # noqa: E402

W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]




def seeded_sequence(seed, T):
    """The canonical injected piece sequence: ONE random.Random(seed), T draws of randrange(7)."""
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


def ce_populations():
    ce = json.load(open(os.path.join(ROOT, "tests", "golden", "ce.json")))["ce"]
    rng = random.Random(ce["rng_seed"])
    pops = []
    mean, std = [0.0] * 6, [ce["stdev"]] * 6
    for g in range(2):
        pops.append(np.array([[rng.gauss(mean[i], std[i]) for i in range(6)] for _ in range(ce["width"])]))
        mean = [float.fromhex(x) for x in ce["generations"][g]["mean"]]
        std = [float.fromhex(x) for x in ce["generations"][g]["std"]]
    seqs = []
    for g in range(ce["games"]):
        r = random.Random(ce["seed0"] + g)
        seqs.append([r.randrange(7) for _ in range(ce["T"])])
    return pops, np.array(seqs, dtype=np.uint8)


def timeit(eng, reps, **kw):
    best = 1e9
    out = None
    for _ in range(reps):
        out = eng.play_games(**kw)
        best = min(best, eng.last_timing()["kernel_ms"])
    return best, out


def main():
    eng = Engine(0)
    eng.set_features(W6_NAMES)
    pops, seqs = ce_populations()
    rows = []
    for gi, pop in enumerate(pops):
        for lanes in (64, 32, 16, 8):
            ms, out = timeit(eng, 5, weights=pop, games_per_candidate=10, pieces=seqs, lanes=lanes)
            pl = int(out["results"]["placements"].sum())
            rows.append(("ce_gen%d 100x10" % gi, lanes, ms, pl, int(out["results"]["pieces"].sum())))
    rng = np.random.default_rng(1)
    for n in (1024, 8192, 65536, 262144):
        w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(n, 6))
        for lanes in ((64, 32, 16, 8) if n <= 8192 else (32, 16, 8)):
            ms, out = timeit(eng, 2, weights=w, games_per_candidate=1, T=400, seed=5, seq_of_game=np.arange(n, dtype=np.int32), lanes=lanes)
            pl = int(out["results"]["placements"].sum())
            rows.append(("sweep %d games T=400" % n, lanes, ms, pl, int(out["results"]["pieces"].sum())))
    # BASELINE configs[2] shard: 10 000 candidates x 30 games over 8 GPUs = 1 250 x 30 per GPU (w6, T = 1000, generation-1-like population)
    import json as _json
    ce = _json.load(open(os.path.join(ROOT, "tests", "golden", "ce.json")))["ce"]
    mean = np.array([float.fromhex(x) for x in ce["generations"][0]["mean"]]); std = np.array([float.fromhex(x) for x in ce["generations"][0]["std"]])
    w3 = mean[None, :] + std[None, :] * rng.normal(0, 1.0, size=(1250, 6))
    seqs30 = np.array([seeded_sequence(1000 + g, 1000) for g in range(30)], dtype=np.uint8)
    for lanes in (0, 32, 8, 4):
        ms, out = timeit(eng, 2, weights=w3, games_per_candidate=30, pieces=seqs30, lanes=lanes)
        rows.append(("config3 shard 1250x30 T=1000", eng.last_timing()["lanes"], ms, int(out["results"]["placements"].sum()), int(out["results"]["pieces"].sum())))
    # all-45 and lookahead-2
    eng.set_features(list(range(45)))
    w = rng.normal(0, 0.3, size=(4096, 45))
    for i, n in enumerate(W6_NAMES):
        from tetris_ai_b200.tetris.features import FEATURE_ID
        w[:, FEATURE_ID[n]] += W6_VALS[i]
    for lanes in (32, 8):
        ms, out = timeit(eng, 2, weights=w, games_per_candidate=1, T=300, seed=5, seq_of_game=np.arange(4096, dtype=np.int32), lanes=lanes)
        rows.append(("all45 L1 4096 games T=300", lanes, ms, int(out["results"]["placements"].sum()), int(out["results"]["pieces"].sum())))
    for lanes in (32, 8):
        ms, out = timeit(eng, 1, weights=w[:512], games_per_candidate=1, T=60, seed=5, seq_of_game=np.arange(512, dtype=np.int32), lanes=lanes, lookahead=2)
        rows.append(("all45 L2 512 games T=60", lanes, ms, int(out["results"]["placements"].sum()), int(out["results"]["pieces"].sum())))
    eng.set_features(W6_NAMES)
    w6 = np.array(W6_VALS)[None, :] + rng.normal(0, 1.0, size=(2048, 6))
    for lanes in (32, 8):
        ms, out = timeit(eng, 1, weights=w6, games_per_candidate=1, T=100, seed=5, seq_of_game=np.arange(2048, dtype=np.int32), lanes=lanes, lookahead=2)
        rows.append(("w6 L2 2048 games T=100", lanes, ms, int(out["results"]["placements"].sum()), int(out["results"]["pieces"].sum())))
    # overhang-slide move set (generic kernels)
    hole_names = ["max_ht", "row_trans", "pits", "landing_height"]
    eng.set_features(hole_names)
    wh = np.array([-1.0, -0.2, 0.4, -0.5])[None, :] + rng.normal(0, 0.2, size=(8192, 4))
    for ms in (0, 1):
        for lanes in (32, 8):
            ms_t, out = timeit(eng, 2, weights=wh, games_per_candidate=1, T=200, seed=5, seq_of_game=np.arange(8192, dtype=np.int32), lanes=lanes, move_set=ms)
            rows.append(("holes4 %s 8192 games T=200" % ("slide" if ms else "drop"), lanes, ms_t, int(out["results"]["placements"].sum()), int(out["results"]["pieces"].sum())))
    eng.set_features(W6_NAMES)
    for lanes in (32, 8):
        ms_t, out = timeit(eng, 2, weights=w6, games_per_candidate=1, T=300, seed=5, seq_of_game=np.arange(2048, dtype=np.int32), lanes=lanes, move_set=1)
        rows.append(("w6 slide 2048 games T=300", lanes, ms_t, int(out["results"]["placements"].sum()), int(out["results"]["pieces"].sum())))
    # wiggle move sets (one warp per game, serial flood fill per rotation)
    for ms, nm in ((4, "wiggle_top"), (8, "wiggle_drop")):
        ms_t, out = timeit(eng, 2, weights=w6, games_per_candidate=1, T=300, seed=5, seq_of_game=np.arange(2048, dtype=np.int32), move_set=ms)
        rows.append(("w6 %s 2048 games T=300" % nm, 32, ms_t, int(out["results"]["placements"].sum()), int(out["results"]["pieces"].sum())))
    print("%-32s %5s %10s %12s %10s %14s" % ("workload", "lanes", "kernel_ms", "placements", "pieces", "placements/s"))
    for name, lanes, ms, pl, pc in rows:
        print("%-32s %5d %10.3f %12d %10d %14.4e" % (name, lanes, ms, pl, pc, pl / (ms * 1e-3)))


if __name__ == "__main__":
    main()
