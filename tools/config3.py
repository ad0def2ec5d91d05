#!/usr/bin/env python3
"""BASELINE configs[2]: CE generations of 10 000 candidates x 30 games (w6, lookahead 1, seeds 1000..1029, T = 1000),
the population sharded by candidate over the ranks, per-candidate fitness all-gathered over NCCL.

    python tools/config3.py                                                        # 1 GPU (all 10 000 candidates on it)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 tools/config3.py
Rank 0 prints one JSON line per generation: wall time of the whole generation (host driver + GPU + all-gather), GPU kernel
time of the local shard, placements, and a digest of the refitted (mean, std) — identical for every world size.
"""
import hashlib
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]




def seeded_sequence(seed, T):
    """The canonical injected piece sequence: ONE random.Random(seed), T draws of randrange(7)."""
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


def main():
    real_stdout = os.dup(1)          # NCCL_DEBUG=INFO writes to fd 1: keep stdout for the JSON lines, send the rest to stderr
    os.dup2(2, 1)
    import torch
    from tetris_ai_b200 import Engine
    from tetris_ai_b200.learning import PopulationEvaluator, ShardedEvaluator, cross_entropy
    from tetris_ai_b200.tetris import features as F
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    width, keep, G, T = int(os.environ.get("CE_WIDTH", "10000")), int(os.environ.get("CE_KEEP", "1000")), 30, 1000
    feats = [getattr(F, n) for n in W6_NAMES]
    seqs = np.array([seeded_sequence(1000 + g, T) for g in range(G)], dtype=np.uint8)
    eng = Engine(local)
    local_ev = PopulationEvaluator(eng, feats, G, pieces=seqs)
    stats = {"kernel_ms": 0.0, "placements": 0}

    raw_evaluate = local_ev.evaluate

    def local_eval(w):
        fit = raw_evaluate(w)
        stats["placements"] = int(local_ev.last["results"]["placements"].sum())
        return fit
    local_ev.evaluate = local_eval
    ev = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = os.environ.get("TB_NCCL_DEBUG", "INFO")      # ranks / transports visible in the log (stderr)
        dist.init_process_group("nccl", device_id=dev)
        ev = ShardedEvaluator(local_ev, device=dev)            # device path: play -> all-gather of the device fitness -> one D2H
        ev.collect_timing = True
        ev.collect_placements = True
        test_f, map_f = ev.test_f, ev.map_f
    else:
        test_f, map_f = local_ev.test_f, local_ev.map_f
    times = {}
    orig_map = map_f

    def timed_map(fn, it):
        t0 = time.perf_counter()
        out = orig_map(fn, it)
        times["map_s"] = time.perf_counter() - t0
        return out
    gen = cross_entropy(feats, 10.0, width, keep, random.Random(42), test_f, map_f=timed_map)
    for g in range(int(os.environ.get("CE_GENERATIONS", "3"))):
        t0 = time.perf_counter()
        mean, std = next(gen)
        wall = time.perf_counter() - t0
        placements = stats["placements"]
        if ev is not None:
            placements = ev.last_timing["placements"]
        if world > 1:
            tot = torch.tensor([float(placements)], dtype=torch.float64, device=dev)
            dist.all_reduce(tot)
            placements = int(tot.item())
        digest = hashlib.sha256(("".join(float(v).hex() for v in list(mean.values()) + list(std.values()))).encode()).hexdigest()[:16]
        if rank == 0:
            os.write(real_stdout, (json.dumps({"generation": g, "n_gpus": world, "width": width, "games_per_candidate": G, "wall_s": wall, "evaluate_s": times["map_s"],
                              "sharded_timing": ({k: round(v, 3) for k, v in ev.last_timing.items()} if ev is not None else None),
                              "placements": placements, "placements_per_s_wall": placements / wall, "placements_per_s_evaluate": placements / times["map_s"],
                              "refit_digest": digest, "mean": [float(v) for v in mean.values()]}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
