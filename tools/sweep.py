#!/usr/bin/env python3
"""BASELINE configs[4]: throughput sweep, 1K .. 1M concurrent games at 1 / 2 / 4 / 8 GPUs.

Every game has its own candidate (w6 + N(0, 1) noise, numpy default_rng(99)) and its own counter-based piece stream
(K4: piece = hash(seed, game, t)); the cap on the game length is 10^6 pieces, lowered per point so that one point stays
around 4e10 placements (stated in the output).  Games are split evenly over the ranks (no data-path collective; one
all-reduce of the placement counters at the end of a point).  At every point 64 random games are replayed through the
plain-C oracle, truncated at 2000 pieces, and compared move by move (bit-exact) with a traced GPU run of the same games.

    python tools/sweep.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 tools/sweep.py
Prints one JSON line per point on rank 0.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = np.array([-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626])
SEED = 2026


def main():
    import torch
    from tetris_ai_b200 import Engine, _lib
    from oracle import coracle, pyoracle
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local); eng.set_features(W6_NAMES)
    int_peak = json.load(open(os.path.join(ROOT, "profiles", "INT_PEAK_r01.json")))["int_peak_thread_ops_per_s"]
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    points = [int(x) for x in os.environ.get("SWEEP_POINTS", "1024,4096,16384,65536,262144,1048576").split(",")]
    for n in points:
        cap = int(min(1_000_000, max(200, 4e10 / (23.0 * n))))
        rng = np.random.default_rng(99)
        w_all = W6_VALS[None, :] + rng.normal(0, 1.0, size=(n, 6))
        lo, hi = rank * n // world, (rank + 1) * n // world
        m = hi - lo
        d_w = torch.from_numpy(np.ascontiguousarray(w_all[lo:hi])).to(dev)
        d_seq = torch.arange(lo, hi, dtype=torch.int32, device=dev)
        d_res = torch.zeros(m * _lib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream()

        def launch():
            eng.play_games_device(st.cuda_stream, m, 1, d_w.data_ptr(), d_res.data_ptr(), cap, seq_of_game_ptr=d_seq.data_ptr(), seed=SEED)
        launch(); torch.cuda.synchronize()                      # warm-up (same work; results are deterministic)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); launch(); e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        lanes_used = eng.last_timing()["lanes"]
        res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=_lib.RESULT_DTYPE)
        stats = torch.tensor([ms, float(res["placements"].sum()), float(res["pieces"].sum()), float(res["pieces"].max())], dtype=torch.float64, device=dev)
        if world > 1:
            mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            ms, placements, pieces, longest = float(mx[0]), float(sm[1]), float(sm[2]), float(mx[3])
        else:
            placements, pieces, longest = float(stats[1]), float(stats[2]), float(stats[3])
        # spot check on this rank's shard: 64 random games, truncated at 2000 pieces, against the plain-C oracle
        pick = np.sort(np.random.default_rng(n + rank).choice(m, size=min(64, m), replace=False))
        tcap = min(cap, 2000)
        g = eng.play_games(w_all[lo:hi][pick], 1, T=tcap, seed=SEED, seq_of_game=(lo + pick).astype(np.int32), want_trace=True)
        o = coracle.play_games(fids, w_all[lo:hi][pick], 1, T=tcap, seed=SEED, seq_of_game=(lo + pick).astype(np.int32), threads=os.cpu_count() // max(world, 1) or 1, want_trace=True)
        ok = bool((g["results"]["pieces"] == o["pieces"]).all() and (g["results"]["score"] == o["score"]).all())
        for i in range(len(pick)):
            k = int(o["pieces"][i])
            ok = ok and bool((g["trace"][i, :k] == o["trace"][i, :k]).all())
            # the full-length run starts with the same moves: its game is at least as long as the truncated one
            ok = ok and int(res["pieces"][pick[i]]) >= k
        flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            rate = placements / (ms * 1e-3)
            print(json.dumps({"games": n, "n_gpus": world, "cap_pieces": cap, "lanes_per_game": lanes_used, "ms": ms, "placements": placements, "pieces": pieces,
                              "longest_game": longest, "placements_per_s": rate, "roofline_frac_int_issue": rate * 220.0 / (int_peak * world),
                              "oracle_spot_check_games": int(len(pick) * world), "oracle_spot_check_ok": bool(flag.item() == 1.0)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
