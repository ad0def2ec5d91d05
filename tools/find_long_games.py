#!/usr/bin/env python3
"""Find counter-based piece streams on which the w6 weights survive T pieces (C oracle), for the long-game parity tests.
Writes tests/golden/long_streams.json = {"seed", "T", "streams": [...]}.   Usage: python tools/find_long_games.py [T] [want]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import coracle, pyoracle      # noqa: E402

W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    seed = 2024
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    found, first = [], 0
    threads = os.cpu_count() or 1
    while len(found) < want:
        ids = np.arange(first, first + threads, dtype=np.int32)
        o = coracle.play_games(fids, np.array([W6_VALS] * len(ids)), 1, T=T, seed=seed, seq_of_game=ids, threads=threads)
        for i, n in zip(ids, o["pieces"]):
            print("stream %d: %d pieces, %d lines" % (i, n, o["lines"][list(ids).index(i)]), flush=True)
            if int(n) == T:
                found.append(int(i))
        first += threads
    json.dump({"seed": seed, "T": T, "streams": found[:max(want, len(found))], "note": "w6 weights survive T pieces on these streams (tools/find_long_games.py, C oracle)"},
              open(os.path.join(ROOT, "tests", "golden", "long_streams.json"), "w"))
    print("streams:", found)


if __name__ == "__main__":
    main()
