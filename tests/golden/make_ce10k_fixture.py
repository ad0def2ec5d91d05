#!/usr/bin/env python3
"""Fixture for BASELINE configs[2]: generation 0 of the 10 000 x 30 cross-entropy run
(w6 features, sigma 10, random.Random(42), piece seeds 1000..1029, T = 1000), evaluated by the plain-C oracle
(oracle/coracle.c — itself pinned to the reference's golden vectors by tests/test_oracle_golden.py).

Writes tests/golden/ce10k_gen0.json.gz: per-candidate total lines (u32[10000]; fitness = lines / 30), total pieces and
placements, and the refit (mean, std as hex doubles) the host driver must produce from it — the anchor for the GPU test
of the full population and for the refit digests of tools/config3.py / bench.py's population_sharded.

~7.9e8 placements: about 5 minutes on 8 cores.   Usage:  python tests/golden/make_ce10k_fixture.py [threads]
"""
import gzip
import hashlib
import heapq
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import coracle, pyoracle          # noqa: E402

W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
WIDTH, KEEP, G, T, SIGMA, RNG_SEED, SEED0 = 10000, 1000, 30, 1000, 10.0, 42, 1000




def seeded_sequence(seed, T):
    """The canonical injected piece sequence: ONE random.Random(seed), T draws of randrange(7)."""
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


def population():
    rng = random.Random(RNG_SEED)
    return np.array([[rng.gauss(0, SIGMA) for _ in range(6)] for _ in range(WIDTH)], dtype=np.float64)


def sequences():
    return np.array([seeded_sequence(SEED0 + g, T) for g in range(G)], dtype=np.uint8)


def refit(pop, fitness):
    scored = [(fitness[i], i) for i in range(len(pop))]
    elite = [pop[i] for _, i in heapq.nlargest(KEEP, scored)]
    mean = [np.mean([e[d] for e in elite]) for d in range(6)]
    std = [np.std([e[d] for e in elite]) for d in range(6)]
    return mean, std


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    pop, seqs = population(), sequences()
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    t0 = time.time()
    o = coracle.play_games(fids, pop, G, pieces=seqs, threads=threads)
    dt = time.time() - t0
    lines = o["lines"].reshape(WIDTH, G).sum(axis=1).astype(np.int64)
    fitness = (lines / G).tolist()
    mean, std = refit(pop.tolist(), fitness)
    digest = hashlib.sha256(("".join(float(v).hex() for v in list(mean) + list(std))).encode()).hexdigest()[:16]
    data = {"meta": {"generator": "tests/golden/make_ce10k_fixture.py", "evaluator": "oracle/coracle.c", "seconds": dt, "threads": threads},
            "width": WIDTH, "keep": KEEP, "games": G, "T": T, "stdev": SIGMA, "rng_seed": RNG_SEED, "seed0": SEED0, "features": W6_NAMES,
            "candidate_lines": [int(x) for x in lines], "pieces": int(o["pieces"].sum()), "placements": int(o["placements"].sum()),
            "score_sum": int(o["score"].sum()),
            "mean": [float(v).hex() for v in mean], "std": [float(v).hex() for v in std], "refit_digest": digest}
    with gzip.open(os.path.join(HERE, "ce10k_gen0.json.gz"), "wt") as f:
        json.dump(data, f)
    print("done in %.0f s: %d placements, %d pieces, digest %s" % (dt, data["placements"], data["pieces"], digest))


if __name__ == "__main__":
    main()
