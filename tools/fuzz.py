#!/usr/bin/env python3
"""Differential fuzzing of the CUDA path against the plain-C oracle: random feature lists / weights / lane widths /
lookahead / move sets / pressure parameters / initial boards / sequence lengths; every output field and every move is
compared.  usage: fuzz.py [iterations] [seed]   (FUZZ_MOVE_SETS=wiggle_top,wiggle_drop restricts the move sets)"""
import os, sys, random, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine
from oracle import coracle, pyoracle

NAMES = list(pyoracle.FEATURE_NAMES)
W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]


def random_board(rng):
    hmax = rng.choice([0, 0, 3, 8, 14, 18, 19])
    rows = [0] * 20
    dens = rng.choice([0.5, 0.75, 0.9])
    for r in range(20 - hmax, 20):
        v = 0
        for c in range(10):
            if rng.random() < dens:
                v |= 1 << c
        rows[r] = v if v != 0x3FF else v & ~(1 << rng.randrange(10))
    return rows


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    # several ctxs with different scheduling switches (read at tb200_create): the lockstep kernel's per-bucket width rule is pushed
    # to every width, the quad kernel takes over the sparse stages, the window evaluation is switched off
    engines = []
    for env in ({}, {"TB200_LOCK_THR": "0,0,0"}, {"TB200_LOCK_THR": "1000000,0,0"}, {"TB200_LOCK_THR": "1000000,1000000,0"},
                {"TB200_LOCK_THR": "1000000,1000000,1000000"}, {"TB200_LOCK_THR": "60,20,6"}, {"TB200_QUAD": "1"}, {"TB200_NO_WIN": "1"}):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        engines.append(Engine(0))
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v
    t0 = time.time()
    stats = {"games": 0, "pieces": 0, "placements": 0}
    for it in range(iters):
        kind = rng.randrange(5)
        if kind == 0:
            names = W6
        elif kind == 1:
            names = NAMES
        else:
            names = rng.sample(NAMES, rng.randint(1, 14))
        ms_name = rng.choice(os.environ.get("FUZZ_MOVE_SETS", "drop,drop,drop,slide,wiggle_top,wiggle_drop").split(","))
        la = 2 if rng.random() < 0.25 else 1
        pressure = None
        if rng.random() < 0.3:
            pressure = {"avg_lat": rng.uniform(100, 350), "resp_time": rng.uniform(30, 130), "move_eff": rng.uniform(0.9, 1.6), "two_but_rot": rng.random() < 0.5}
        lanes = rng.choice([0, 128, 64, 32, 16, 8, 4, 2, 1])
        eng = rng.choice(engines)
        C, G = rng.randint(1, 24), rng.randint(1, 3)
        T = rng.randint(1, 30) if (la == 2 or ms_name.startswith("wiggle")) else rng.randint(1, 220)
        if la == 2 and ms_name.startswith("wiggle"):
            T = rng.randint(1, 10); C = min(C, 4)
        if rng.random() < float(os.environ.get("FUZZ_STAGED", "0.03")):
            # latency regime: automatic lanes, more than 2 games per SM, >= 128 pieces -> the re-spreading stages run
            names, ms_name, la, pressure, lanes = W6, "drop", 1, None, 0
            C, G, T = rng.randint(300, 700), 1, rng.randint(128, 200)
        if rng.random() < float(os.environ.get("FUZZ_LOCKSTEP", "0.02")):
            # population regime: > 80 games per SM on shared sequences -> the lockstep kernel K2s
            names, ms_name, la, pressure, lanes = (W6 if rng.random() < 0.5 else names), "drop", 1, None, 0      # K2s exists for every feature mode
            G = rng.randint(4, 12); C = 12200 // G + rng.randint(0, 300); T = rng.randint(128, 150)
        binned = rng.random() < float(os.environ.get("FUZZ_BINNED", "0.01"))
        if binned:
            # independent streams, >= 32 768 games -> the piece-binned kernel K2b (per-game sequences come from seq_of_game = g % G with G = n)
            names, ms_name, la, pressure, lanes = W6, "drop", 1, None, 0
            C, G, T = 1, rng.randint(32768, 36000), rng.randint(8, 20)
        if names is W6 and rng.random() < 0.6:
            w = np.array([[rng.gauss(v, 2.0) for v in W6V] for _ in range(C)])
        elif rng.random() < 0.2:
            w = np.array([[float(rng.randint(-2, 2)) for _ in names] for _ in range(C)])      # many ties
        else:
            w = np.array([[rng.gauss(0, 3.0) for _ in names] for _ in range(C)])
        n = C * G
        use_pieces = rng.random() < 0.5
        pieces = np.array([[rng.randrange(7) for _ in range(T)] for _ in range(G)], dtype=np.uint8) if use_pieces else None
        init = None
        if rng.random() < 0.4 and n <= 2000:
            init = np.array([random_board(rng) for _ in range(n)], dtype=np.uint16)
        fids = [pyoracle.FEATURE_ID[x] for x in names]
        bit = {"drop": 0, "slide": 1, "wiggle_top": 4, "wiggle_drop": 8}[ms_name]
        seed = rng.getrandbits(40)
        eng.set_features(names)
        r = eng.play_games(w, G, pieces=pieces, T=T, seed=seed, lookahead=la, lanes=lanes, init_rows=init, want_trace=True, want_final=True,
                           move_set=bit, pressure=pressure)
        res = r["results"]
        if init is None and n > 64:
            # big cases (staged / lockstep / binned populations): the oracle's multi-threaded batch entry, whole arrays compared
            o = coracle.play_games(fids, w, G, pieces=pieces, T=T, seed=seed, lookahead=la, threads=os.cpu_count() or 8, want_trace=True,
                                   move_set=ms_name, pressure=pressure)
            ok = ((res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all() and (res["hist"] == o["hist"]).all()
                  and (res["score"] == o["score"]).all() and (res["placements"] == o["placements"]).all())
            if ok:
                played = np.arange(T)[None, :, None] < o["pieces"].astype(np.int64)[:, None, None]
                ok = bool((np.where(played, r["trace"], 0) == np.where(played, o["trace"], 0)).all())
            if not ok:
                print("MISMATCH iteration", it, dict(names=names, ms=ms_name, la=la, pressure=pressure, lanes=lanes, C=C, G=G, T=T, seed=seed, use_pieces=use_pieces))
                sys.exit(1)
            stats["games"] += n; stats["pieces"] += int(o["pieces"].sum()); stats["placements"] += int(o["placements"].sum())
            continue
        for g in range(n):
            seq = pieces[g % G] if use_pieces else [pyoracle.hashed_piece(seed, g % G, t) for t in range(T)]
            o = coracle.play_one(fids, w[g // G], seq, la, init_rows=None if init is None else init[g], move_set=ms_name, pressure=pressure)
            k = o["pieces"]
            ok = (int(res["pieces"][g]) == k and int(res["lines"][g]) == o["lines"] and [int(x) for x in res["hist"][g]] == o["hist"]
                  and int(res["score"][g]) == o["score"] and int(res["placements"][g]) == o["placements"]
                  and r["trace"][g, :k].tobytes() == o["trace"].tobytes() and list(r["final"][g]) == list(o["final"]))
            if not ok:
                print("MISMATCH iteration", it, "game", g, dict(names=names, ms=ms_name, la=la, pressure=pressure, lanes=lanes, C=C, G=G, T=T, seed=seed,
                                                              use_pieces=use_pieces, init=init is not None))
                print(" gpu:", res[g], r["trace"][g, :k + 1].tolist()[:12])
                print(" cpu:", o["pieces"], o["lines"], o["hist"], o["score"], o["placements"], o["trace"].tolist()[:12])
                sys.exit(1)
            stats["games"] += 1; stats["pieces"] += k; stats["placements"] += o["placements"]
        if time.time() - t0 > float(os.environ.get("FUZZ_SECONDS", "1e9")):
            iters = it + 1
            break
    print("fuzz ok: %d iterations, %d games, %d pieces, %d placements, %.1f s" % (iters, stats["games"], stats["pieces"], stats["placements"], time.time() - t0))


if __name__ == "__main__":
    main()
