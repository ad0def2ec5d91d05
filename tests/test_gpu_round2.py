"""Round-2 parity tests on a real GPU, through the C ABI (VERDICT r1 "what's weak" 1-4, "missing" 3):

  * BASELINE configs[3] exactly as SURVEY.md §8d states it (F = 45, Random(7).gauss(0,1), lookahead 2, seeds 0..9, T = 200)
    and a 45-feature lookahead-2 weight set that survives all 200 pieces, against the C oracle;
  * one >= 200 000-piece game per kernel family on the counter-based piece stream against coracle.play_one;
  * generation 0 of the 10 000 x 30 population (BASELINE configs[2], 7.9e8 placements) against the committed C-oracle
    fixture, with a live oracle re-check of a random subset of candidates;
  * State.score(points) at the C ABI (tb200_play_args.points) for every kernel family;
  * one call in flight per ctx: device-mode calls on different streams are ordered by the library;
  * refusal of non-finite weights; tb200_check_status.
Bit-exact everywhere.
"""
import gzip
import hashlib
import json
import os
import random

import numpy as np
import pytest

from conftest import load_golden
from oracle import coracle, pyoracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NAMES = list(pyoracle.FEATURE_NAMES)
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
THREADS = os.cpu_count() or 8


@pytest.fixture(scope="module")
def eng():
    from tetris_ai_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def same_results(res, o):
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all()
    assert (res["hist"] == o["hist"]).all() and (res["score"] == o["score"]).all()
    assert (res["placements"] == o["placements"]).all()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[3] = SURVEY.md §8d config 4
# ---------------------------------------------------------------------------------------------------------------------
def all45_survivor_weights():
    """w6's values on its six features, small seeded noise on the other 39: plays like w6, so lookahead-2 games reach the cap."""
    rng = random.Random(11)
    return [(W6_VALS[W6_NAMES.index(n)] if n in W6_NAMES else 0.0) + rng.gauss(0, 0.02) for n in NAMES]


@pytest.mark.parametrize("which", ["survey", "survivor"])
def test_config4_two_piece_lookahead_all45_exact(eng, which):
    rng = random.Random(7)
    w = [rng.gauss(0, 1) for _ in range(45)] if which == "survey" else all45_survivor_weights()
    T = 200
    seqs = np.array([pyoracle.piece_sequence(seed, T) for seed in range(10)], dtype=np.uint8)
    fids = list(range(45))
    eng.set_features(NAMES)
    o = coracle.play_games(fids, np.array([w]), 10, pieces=seqs, lookahead=2, threads=THREADS, want_trace=True)
    if which == "survivor":
        assert (o["pieces"] == T).all(), o["pieces"]          # the point of this weight set: the K2L path runs the full 200 pieces
        assert int(o["placements"].sum()) > 500000
    for lanes in (0, 32, 16):
        r = eng.play_games(np.array([w]), 10, pieces=seqs, lookahead=2, lanes=lanes, want_trace=True, want_scores=True)
        same_results(r["results"], o)
        for g in range(10):
            n = int(o["pieces"][g])
            assert r["trace"][g, :n].tobytes() == o["trace"][g, :n].tobytes(), (which, lanes, g)


# ---------------------------------------------------------------------------------------------------------------------
# long games (BASELINE configs[4] caps at 10^6 pieces; r1 never played more than 1 000)
# ---------------------------------------------------------------------------------------------------------------------
LONG_T = 200000
LONG_SEED = 2024


def strong_streams(n):
    """Stream ids whose w6 game survives LONG_T pieces (found with the C oracle by tools/find_long_games.py)."""
    ids = json.load(open(os.path.join(ROOT, "tests", "golden", "long_streams.json")))
    assert ids["seed"] == LONG_SEED and ids["T"] == LONG_T
    return ids["streams"][:n]


@pytest.fixture(scope="module")
def long_oracle():
    """C-oracle results of the long w6 games, computed once (threads in parallel): {stream: result dict}."""
    streams = strong_streams(4)
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    o = coracle.play_games(fids, np.array([W6_VALS] * len(streams)), 1, T=LONG_T, seed=LONG_SEED,
                           seq_of_game=np.array(streams, dtype=np.int32), threads=len(streams), want_trace=True)
    out = {}
    for i, s in enumerate(streams):
        n = int(o["pieces"][i])
        out[s] = {"pieces": n, "lines": int(o["lines"][i]), "score": int(o["score"][i]), "placements": int(o["placements"][i]),
                  "hist": [int(x) for x in o["hist"][i]], "trace_sha": hashlib.sha256(o["trace"][i, :n].tobytes()).hexdigest()}
    assert all(v["pieces"] == LONG_T for v in out.values())
    return out


def check_long(res, trace, want):
    assert int(res["pieces"]) == want["pieces"] and int(res["lines"]) == want["lines"]
    assert int(res["score"]) == want["score"] and int(res["placements"]) == want["placements"]
    assert [int(x) for x in res["hist"]] == want["hist"]
    if trace is not None:
        assert hashlib.sha256(trace[: want["pieces"]].tobytes()).hexdigest() == want["trace_sha"]


@pytest.mark.parametrize("lanes,family", [(32, "K2f one warp per game"), (64, "K2p two warps per game"), (128, "K2q four warps per game"), (8, "K2g 8 lanes"), (1, "K2g 1 lane")])
def test_long_game_single_launch_families(eng, long_oracle, lanes, family):
    streams = list(long_oracle)[:2]
    eng.set_features(W6_NAMES)
    r = eng.play_games(np.array([W6_VALS] * len(streams)), 1, T=LONG_T, seed=LONG_SEED, seq_of_game=np.array(streams, dtype=np.int32),
                       lanes=lanes, want_trace=True)
    for i, s in enumerate(streams):
        check_long(r["results"][i], r["trace"][i], long_oracle[s])


def test_long_game_generic_feature_list(eng, long_oracle):
    """The run-time feature list code (MODE_GENERIC): w6 plus a zero-weight seventh feature plays the same game."""
    streams = list(long_oracle)[:1]
    names = W6_NAMES + ["max_ht"]
    eng.set_features(names)
    for lanes in (32, 16):
        r = eng.play_games(np.array([W6_VALS + [0.0]]), 1, T=LONG_T, seed=LONG_SEED, seq_of_game=np.array(streams, dtype=np.int32),
                           lanes=lanes, want_trace=True)
        check_long(r["results"][0], r["trace"][0], long_oracle[streams[0]])


def _weak_weights(rng, n):
    """Anti-w6 candidates (every weight has the wrong sign): they reward holes and height, so their games end within a few
    dozen pieces whatever the stream — the oracle can check all of them, and only the planted strong games run long."""
    return -np.sign(np.array(W6_VALS))[None, :] * np.abs(rng.normal(0, 10.0, size=(n, 6)))


def test_long_game_staged_latency_path(eng, long_oracle):
    """Re-spreading stages (2..16 games per SM): a thousand weak games die early, two strong ones run 200 000 pieces through
    every stage hand-over (K2f dense -> K2f sparse -> K2p -> the quad stage)."""
    streams = list(long_oracle)[:2]
    rng = np.random.default_rng(3)
    n = 1000
    w = _weak_weights(rng, n)
    seqid = np.arange(n, dtype=np.int32) + 100000
    for i, s in enumerate(streams):
        w[17 + 400 * i] = W6_VALS
        seqid[17 + 400 * i] = s
    eng.set_features(W6_NAMES)
    r = eng.play_games(w, 1, T=LONG_T, seed=LONG_SEED, seq_of_game=seqid)
    assert eng.last_timing()["kernels_launched"] >= 3
    for i, s in enumerate(streams):
        check_long(r["results"][17 + 400 * i], None, long_oracle[s])
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    short = [g for g in range(n) if g not in (17, 417)]
    o = coracle.play_games(fids, w[short], 1, T=3000, seed=LONG_SEED, seq_of_game=seqid[short], threads=THREADS)
    assert (o["pieces"] < 3000).all()
    same_results(r["results"][short], o)


def test_long_game_piece_binned_kernel(eng, long_oracle):
    """K2b (>= 32 768 games on independent streams): the strong games make 200 000 trips through the per-piece queues."""
    streams = list(long_oracle)[:3]
    rng = np.random.default_rng(4)
    n = 33000
    w = _weak_weights(rng, n)
    seqid = np.arange(n, dtype=np.int32) + 100000
    where = [5, 12345, 32999]
    for g, s in zip(where, streams):
        w[g] = W6_VALS
        seqid[g] = s
    eng.set_features(W6_NAMES)
    r = eng.play_games(w, 1, T=LONG_T, seed=LONG_SEED, seq_of_game=seqid)
    for g, s in zip(where, streams):
        check_long(r["results"][g], None, long_oracle[s])
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    short = np.setdiff1d(np.arange(n), where)
    o = coracle.play_games(fids, w[short], 1, T=3000, seed=LONG_SEED, seq_of_game=seqid[short], threads=THREADS)
    assert (o["pieces"] < 3000).all()
    same_results(r["results"][short], o)


def test_long_game_lockstep_kernel(eng, long_oracle):
    """K2s (C x G games on G shared sequences, > 80 games per SM): 3 000 weak candidates and one strong one on four shared
    streams, cap 200 000: the strong candidate's four games walk 3 000+ stages."""
    streams = list(long_oracle)[:4]
    rng = np.random.default_rng(5)
    C, G = 3100, 4
    w = _weak_weights(rng, C)
    w[1234] = W6_VALS
    eng.set_features(W6_NAMES)
    # game g of every candidate plays stream streams[g]: the counter-based stream with an explicit (shared) sequence table
    seqs = np.array([[pyoracle.hashed_piece(LONG_SEED, s, t) for t in range(4000)] for s in streams], dtype=np.uint8)
    pieces = np.empty((G, LONG_T), dtype=np.uint8)
    got = eng.hashed_pieces(LONG_SEED, 0, 1, 8)          # K4 smoke (the full tables come from the library too, below)
    assert list(got[0]) == [pyoracle.hashed_piece(LONG_SEED, 0, t) for t in range(8)]
    for i, s in enumerate(streams):
        pieces[i] = eng.hashed_pieces(LONG_SEED, s, 1, LONG_T)[0]
        assert (pieces[i, :4000] == seqs[i]).all()
    r = eng.play_games(w, G, pieces=pieces)
    res = r["results"].reshape(C, G)
    for i, s in enumerate(streams):
        check_long(res[1234, i], None, long_oracle[s])
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    weak = np.setdiff1d(np.arange(C), [1234])
    o = coracle.play_games(fids, w[weak], G, pieces=pieces[:, :3000], threads=THREADS)
    assert (o["pieces"] < 3000).all()
    same_results(res[weak].reshape(-1), o)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2]: generation 0 of the 10 000 x 30 population, whole population on one GPU
# ---------------------------------------------------------------------------------------------------------------------
def test_ce10k_generation0_matches_oracle_fixture(eng):
    fx = json.load(gzip.open(os.path.join(ROOT, "tests", "golden", "ce10k_gen0.json.gz"), "rt"))
    C, G, T = fx["width"], fx["games"], fx["T"]
    rng = random.Random(fx["rng_seed"])
    pop = np.array([[rng.gauss(0, fx["stdev"]) for _ in range(6)] for _ in range(C)])
    seqs = np.array([pyoracle.piece_sequence(fx["seed0"] + g, T) for g in range(G)], dtype=np.uint8)
    eng.set_features(fx["features"])
    r = eng.play_games(pop, G, pieces=seqs)
    res = r["results"]
    lines = res["lines"].reshape(C, G).sum(axis=1)
    assert [int(x) for x in lines] == fx["candidate_lines"]
    assert int(res["pieces"].sum()) == fx["pieces"] and int(res["placements"].sum()) == fx["placements"]
    assert int(res["score"].sum()) == fx["score_sum"]
    assert (r["fitness"] == lines / G).all()
    # the refit the host driver derives from it (learning.py:41-46) — the digest tools/config3.py and bench.py report
    import heapq
    fit = r["fitness"].tolist()
    elite = [pop[i].tolist() for _, i in heapq.nlargest(fx["keep"], [(fit[i], i) for i in range(C)])]
    mean = [np.mean([e[d] for e in elite]) for d in range(6)]
    std = [np.std([e[d] for e in elite]) for d in range(6)]
    assert [float(v).hex() for v in mean] == fx["mean"] and [float(v).hex() for v in std] == fx["std"]
    # live re-check of the fixture itself: 300 random candidates through the C oracle now
    pick = np.sort(np.random.default_rng(0).choice(C, 300, replace=False))
    fids = [pyoracle.FEATURE_ID[n] for n in fx["features"]]
    o = coracle.play_games(fids, pop[pick], G, pieces=seqs, threads=THREADS)
    same_results(res.reshape(C, G)[pick].reshape(-1), o)


# ---------------------------------------------------------------------------------------------------------------------
# State.score(points) at the C ABI
# ---------------------------------------------------------------------------------------------------------------------
def score_from_trace(trace, pieces, pts, lines0=0):
    """game.py:159-164 restated on a move trace: sum over moves of points[rows cleared] * (level after the move + 1)."""
    k = trace[..., 3].astype(np.int64)
    played = np.arange(trace.shape[1])[None, :] < pieces[:, None]
    lines_after = lines0 + np.cumsum(np.where(played, k, 0), axis=1)
    per_move = np.array(pts, dtype=np.int64)[k] * (lines_after // 10 + 1)
    return np.where(played, per_move, 0).sum(axis=1)


def test_custom_points_golden_games_every_family(eng, golden_games):
    games = golden_games["games"]
    n_checked = 0
    for rec in load_golden("r2_extra.json")["points"]:
        g = games[rec["game"]]
        if g.get("tie") == "last":                 # golden game played with a last-of-max tie-breaker: not a device policy
            continue
        names = [n for n, _ in g["weights"]]
        w = np.array([[x for _, x in g["weights"]]])
        seq = np.array([pyoracle.piece_sequence(g["seed"], g["T"])], dtype=np.uint8)
        eng.set_features(names)
        w6 = names == W6_NAMES and g["lookahead"] == 1
        for lanes in ((128, 64, 32, 16, 8, 4, 1) if w6 else (32, 16, 8)):
            for pts, want in rec["scores"]:
                r = eng.play_games(w, 1, pieces=seq, lookahead=g["lookahead"], lanes=lanes, points=pts)
                assert int(r["results"]["score"][0]) == want, (rec["game"], lanes, pts)
                n_checked += 1
    assert n_checked > 200


@pytest.mark.parametrize("shape", ["staged", "lockstep", "binned"])
def test_custom_points_population_schedulers(eng, shape):
    rng = np.random.default_rng(8)
    pts = (7, 50, 150, 350, 1500)
    eng.set_features(W6_NAMES)
    if shape == "staged":
        C, G, T, kw = 1000, 1, 300, {"seq_of_game": (np.arange(1000, dtype=np.int32) * 3) % 50}
    elif shape == "lockstep":
        C, G, T, kw = 3000, 4, 200, {}
    else:
        C, G, T, kw = 33000, 1, 60, {"seq_of_game": np.arange(33000, dtype=np.int32)}
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 4.0, size=(C, 6))
    r = eng.play_games(w, G, T=T, seed=31, want_trace=True, points=pts, **kw)
    res = r["results"]
    want = score_from_trace(r["trace"], res["pieces"].astype(np.int64), pts)
    assert (res["score"].astype(np.int64) == want).all()
    assert int(want.sum()) > 0
    d = eng.play_games(w, G, T=T, seed=31, **kw)["results"]       # default table: only the score may differ
    for f in ("pieces", "lines", "hist", "placements"):
        assert (d[f] == res[f]).all()
    assert (d["score"].astype(np.int64) == score_from_trace(r["trace"], res["pieces"].astype(np.int64), (0, 40, 100, 300, 1200))).all()


# ---------------------------------------------------------------------------------------------------------------------
# boundary behaviour
# ---------------------------------------------------------------------------------------------------------------------
def test_one_call_in_flight_per_ctx_streams_are_ordered(eng):
    """two TB200_MEM_DEVICE calls on DIFFERENT streams of one ctx share the ctx workspace (tickets, hand-over lists);
    the library orders the second behind the first, so both produce exactly what they produce alone."""
    import torch
    from tetris_ai_b200 import _lib
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(12)
    eng.set_features(W6_NAMES)
    jobs = []
    for n, T, seed in ((1000, 400, 1), (900, 500, 2), (3000 * 4, 150, 3)):
        G = 4 if n == 12000 else 1
        C = n // G
        w = np.array(W6_VALS)[None, :] + rng.normal(0, 5.0, size=(C, 6))
        alone = eng.play_games(w, G, T=T, seed=seed)["results"]
        jobs.append({"C": C, "G": G, "T": T, "seed": seed, "alone": alone, "d_w": torch.from_numpy(w).to(dev),
                     "d_res": torch.zeros(n * _lib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev), "stream": torch.cuda.Stream(dev)})
    torch.cuda.synchronize()
    for rep in range(3):
        for j in jobs:
            j["d_res"].zero_()
        torch.cuda.synchronize()
        for j in jobs:                                   # enqueued back to back, nothing synchronised in between
            eng.play_games_device(j["stream"].cuda_stream, j["C"], j["G"], j["d_w"].data_ptr(), j["d_res"].data_ptr(), j["T"], seed=j["seed"])
        eng.check_status()
        torch.cuda.synchronize()
        for j in jobs:
            got = np.frombuffer(j["d_res"].cpu().numpy().tobytes(), dtype=_lib.RESULT_DTYPE)
            assert (got == j["alone"]).all(), (rep, j["C"])


def test_non_finite_weights_are_refused(eng):
    from tetris_ai_b200 import EngineError
    eng.set_features(W6_NAMES)
    for bad in (float("nan"), float("inf"), -float("inf"), 1e301):
        w = np.array([W6_VALS, W6_VALS])
        w[1, 3] = bad
        with pytest.raises(EngineError):
            eng.play_games(w, 1, T=10, seed=1)
    assert int(eng.play_games(np.array([W6_VALS]), 1, T=10, seed=1)["results"]["pieces"][0]) == 10
    with pytest.raises(EngineError):
        eng.play_games(np.array([W6_VALS]), 1, T=10, seed=1, lookahead=3)       # TB200_MAX_LOOKAHEAD


def test_feature_handle_counters_and_piece_stream(eng):
    """feature.accesses / feature.misses on the device handles (feature.py:27,31,36-37) and PieceStream push-back on the GPU."""
    import collections
    from tetris_ai_b200 import feature
    from tetris_ai_b200.tetris import features as F
    from tetris_ai_b200.tetris import simulator as sim
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    feature.set_default_engine(eng)
    st = State(None, Board(20, 10)).future(PIECES[1], 0, 18, 3)
    a0, m0, b0, n0 = F.pits.accesses, F.pits.misses, F.max_ht.accesses, F.max_ht.misses
    cache = {}
    assert F.pits(st, cache=cache) == 2 and F.max_ht(st, cache=cache) == 2 and F.pits(st, cache=cache) == 2       # T pointing down on an empty board
    assert (F.pits.accesses - a0, F.pits.misses - m0) == (2, 1) and (F.max_ht.accesses - b0, F.max_ht.misses - n0) == (1, 1)
    feature.evaluate(st, [F.pits, F.max_ht])
    assert F.pits.accesses - a0 == 3 and F.pits.misses - m0 == 2
    for case in load_golden("r2_extra.json")["consumption"]:
        src = sim.PieceStream(PIECES[i] for i in pyoracle.piece_sequence(case["seed"], case["total"]))
        policy = sim.policy_best(sim.linear_scorer(collections.OrderedDict((getattr(F, n), w) for n, w in case["weights"])), sim.tie_first)
        gen = sim.simulate(State(None, Board(20, 10)), src, sim.move_drop, policy, lookahead=case["lookahead"], engine=eng, chunk=16)
        moves = 0
        for _ in gen:
            moves += 1
            if case["stop_after"] is not None and moves == case["stop_after"]:
                break
        gen.close()
        assert moves == case["moves"]
        first = next(sim.simulate(State(None, Board(20, 10)), src, sim.move_drop, policy, lookahead=case["lookahead"], engine=eng, chunk=16), None)
        assert (None if first is None else first.delta.zoid.name) == case["second_game_first_piece"], case


# ---------------------------------------------------------------------------------------------------------------------
# K2q: four warps per game, three lanes per placement
# ---------------------------------------------------------------------------------------------------------------------
def test_k2q_golden_games(eng, golden_games):
    n = 0
    for g in golden_games["games"]:
        if g.get("tie") == "last" or g["lookahead"] != 1 or [x for x, _ in g["weights"]] != W6_NAMES:
            continue
        eng.set_features(W6_NAMES)
        w = np.array([[x for _, x in g["weights"]]])
        seq = np.array([pyoracle.piece_sequence(g["seed"], g["T"])], dtype=np.uint8)
        r = eng.play_games(w, 1, pieces=seq, lanes=128, want_trace=True, want_final=True, want_scores=True)
        res = r["results"][0]
        assert int(res["pieces"]) == g["pieces"] and int(res["lines"]) == g["lines"] and [int(x) for x in res["hist"]] == g["hist"]
        assert int(res["score"]) == g["score"] and int(res["placements"]) == g["scorer_calls"]
        assert r["trace"][0, :g["pieces"]].tobytes().hex() == g["trace_hex"]
        assert list(r["final"][0]) == g["final_board"]
        n += 1
    assert n >= 3


@pytest.mark.parametrize("kind,C,G,T,spread", [("w6", 60, 4, 400, 1.0), ("w6", 150, 3, 300, 6.0), ("wide", 200, 2, 300, 10.0)])
def test_k2q_vs_coracle_and_other_kernels(eng, kind, C, G, T, spread):
    """Explicit lanes = 128 (one persistent K2q launch): results, every move, every winning score and the final boards
    equal the C oracle's and the one-warp kernel's, from empty boards and from random initial boards with lines."""
    rng = np.random.default_rng(C + T)
    w = (np.array(W6_VALS)[None, :] + rng.normal(0, spread, size=(C, 6))) if kind == "w6" else rng.normal(0, spread, size=(C, 6))
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    eng.set_features(W6_NAMES)
    r = eng.play_games(w, G, T=T, seed=77, lanes=128, want_trace=True, want_scores=True, want_final=True)
    assert eng.last_timing()["lanes"] == 128
    o = coracle.play_games(fids, w, G, T=T, seed=77, threads=THREADS, want_trace=True)
    same_results(r["results"], o)
    for g in range(C * G):
        k = int(o["pieces"][g])
        assert r["trace"][g, :k].tobytes() == o["trace"][g, :k].tobytes(), g
    b = eng.play_games(w, G, T=T, seed=77, lanes=32, want_trace=True, want_scores=True, want_final=True)
    assert (b["results"] == r["results"]).all() and (b["final"] == r["final"]).all()
    for g in range(C * G):
        k = int(o["pieces"][g])
        assert b["scores"][g, :k].tobytes() == r["scores"][g, :k].tobytes(), g
    # initial boards (with garbage rows, some nearly full) and lines already cleared, ragged max_moves
    n = C * G
    init = np.zeros((n, 20), dtype=np.uint16)
    init[::2, 19] = 0x3FE; init[::3, 18] = 0x1FF; init[::3, 19] |= 0x1FF; init[::7, 17] = 0x2AA; init[::7, 18] |= 0x155
    lines0 = (np.arange(n, dtype=np.uint32) * 11) % 73
    kw = dict(T=T, seed=78, init_rows=init, init_lines=lines0, max_moves=T - 13, want_trace=True, want_final=True, want_scores=True)
    a = eng.play_games(w, G, lanes=128, **kw)
    c = eng.play_games(w, G, lanes=32, **kw)
    assert (a["results"] == c["results"]).all() and (a["final"] == c["final"]).all()
    assert a["trace"].tobytes() == c["trace"].tobytes() and a["scores"].tobytes() == c["scores"].tobytes()


# ---------------------------------------------------------------------------------------------------------------------
# K2sW: the lockstep kernel with window-local evaluation (one lane per game)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng_lane1():
    """An Engine whose lockstep kernel is forced to one lane per game (TB200_LOCK_LANES is read at tb200_create): every
    lockstep launch below runs play_games_w6lockwin_kernel, whatever the population size."""
    from tetris_ai_b200 import Engine
    old = os.environ.get("TB200_LOCK_LANES")
    os.environ["TB200_LOCK_LANES"] = "1"
    try:
        e = Engine(0)
    finally:
        if old is None:
            del os.environ["TB200_LOCK_LANES"]
        else:
            os.environ["TB200_LOCK_LANES"] = old
    yield e
    e.close()


@pytest.mark.parametrize("C,G,T,spread,explicit", [(500, 30, 200, 1.0, True), (3001, 5, 131, 6.0, True), (2000, 7, 260, 2.0, False), (13000, 4, 128, 8.0, True), (420, 30, 1000, 0.7, True)])
def test_k2sw_window_lockstep_vs_coracle_and_per_game_kernel(eng_lane1, C, G, T, spread, explicit):
    eng = eng_lane1
    rng = np.random.default_rng(C + G)
    w = np.array(W6_VALS)[None, :] + rng.normal(0, spread, size=(C, 6))
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    eng.set_features(W6_NAMES)
    pieces = np.array([pyoracle.piece_sequence(7 * g + 1, T) for g in range(G)], dtype=np.uint8) if explicit else None
    kw = dict(pieces=pieces, T=T, seed=77)
    a = eng.play_games(w, G, want_trace=True, want_scores=True, want_final=True, **kw)
    assert eng.last_timing()["kernels_launched"] == 2 and eng.last_timing()["lanes"] == 1
    b = eng.play_games(w, G, lanes=8, want_trace=True, want_scores=True, want_final=True, **kw)
    assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all() and (a["fitness"] == b["fitness"]).all()
    for g in range(0, C * G, 3):
        k = int(a["results"]["pieces"][g])
        assert a["trace"][g, :k].tobytes() == b["trace"][g, :k].tobytes(), g
        assert a["scores"][g, :k].tobytes() == b["scores"][g, :k].tobytes(), g
    n_o = min(C, 400)
    o = coracle.play_games(fids, w[:n_o], G, pieces=pieces, T=T, seed=77, threads=THREADS, want_trace=True)
    res = a["results"][: n_o * G]
    same_results(res, o)
    for g in range(0, n_o * G, 5):
        k = int(res["pieces"][g])
        assert a["trace"][g, :k].tobytes() == o["trace"][g, :k].tobytes(), g


def test_k2sw_initial_boards_lines_points_and_max_moves(eng_lane1):
    eng = eng_lane1
    rng = np.random.default_rng(3)
    C, G, T = 700, 20, 400
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(C, 6))
    eng.set_features(W6_NAMES)
    n = C * G
    init = np.zeros((n, 20), dtype=np.uint16)
    init[::3, 19] = 0x1EF; init[::5, 18] = 0x3F0; init[::5, 19] |= 0x3F0; init[::11, 15] = 0x201; init[::11, 16] = 0x3FE; init[::11, 17] = 0x1FF
    lines0 = (np.arange(n, dtype=np.uint32) * 13) % 57
    kw = dict(T=T, seed=5, init_rows=init, init_lines=lines0, max_moves=150, want_trace=True, want_final=True, want_scores=True, points=(3, 40, 100, 300, 1200))
    a = eng.play_games(w, G, **kw)
    assert eng.last_timing()["kernels_launched"] == 2 and eng.last_timing()["lanes"] == 1
    b = eng.play_games(w, G, lanes=32, **kw)
    assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all()
    assert a["trace"].tobytes() == b["trace"].tobytes() and a["scores"].tobytes() == b["scores"].tobytes()


def _engine_with_env(**env):
    from tetris_ai_b200 import Engine
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return Engine(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


# ---------------------------------------------------------------------------------------------------------------------
# K2sP: two warps share the slots of 32 games (opt-in experiment for shards that do not fill the machine: TB200_LOCK_PAIR)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng_pair():
    """An Engine whose per-bucket-width kernel plays EVERY bucket that has at least one game per resident warp with pairs of
    warps (and the thinner ones with a warp per game), so that the pair task, the hand-over between pair and plain buckets and
    the pairs' shared ticket are exercised whatever the population size."""
    e = _engine_with_env(TB200_LOCK_THR="1000000,1,1", TB200_LOCK_PAIR="1")
    yield e
    e.close()


@pytest.mark.parametrize("C,G,T,spread,explicit", [(3001, 5, 131, 6.0, True), (2000, 7, 260, 2.0, False), (420, 30, 400, 0.7, True)])
def test_k2sp_pairs_of_warps_vs_coracle_and_per_game_kernel(eng_pair, C, G, T, spread, explicit):
    """simulate() for a CE population (simulator.py:152-190, learning.py:38) played by pairs of warps that split each piece's slots:
    results, traces, winning scores and final boards equal the per-game kernel's and the C oracle's, scorer-call counts included
    (each warp counts its own share)."""
    eng = eng_pair
    rng = np.random.default_rng(C + G)
    w = np.array(W6_VALS)[None, :] + rng.normal(0, spread, size=(C, 6))
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    eng.set_features(W6_NAMES)
    pieces = np.array([pyoracle.piece_sequence(7 * g + 1, T) for g in range(G)], dtype=np.uint8) if explicit else None
    kw = dict(pieces=pieces, T=T, seed=77)
    a = eng.play_games(w, G, want_trace=True, want_scores=True, want_final=True, **kw)
    assert eng.last_timing()["kernels_launched"] == 2 and eng.last_timing()["lanes"] == 2, eng.last_timing()
    b = eng.play_games(w, G, lanes=8, want_trace=True, want_scores=True, want_final=True, **kw)
    assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all() and (a["fitness"] == b["fitness"]).all()
    for g in range(0, C * G, 3):
        k = int(a["results"]["pieces"][g])
        assert a["trace"][g, :k].tobytes() == b["trace"][g, :k].tobytes(), g
        assert a["scores"][g, :k].tobytes() == b["scores"][g, :k].tobytes(), g
    n_o = min(C, 300)
    o = coracle.play_games(fids, w[:n_o], G, pieces=pieces, T=T, seed=77, threads=THREADS, want_trace=True)
    res = a["results"][: n_o * G]
    same_results(res, o)
    for g in range(0, n_o * G, 5):
        k = int(res["pieces"][g])
        assert a["trace"][g, :k].tobytes() == o["trace"][g, :k].tobytes(), g
    # repeatable (the pairs meet at an mbarrier every piece; tickets are taken by one warp for both)
    a2 = eng.play_games(w, G, **kw)
    assert (a2["results"] == a["results"]).all()


def test_k2sp_initial_boards_lines_points_max_moves_and_pressure(eng_pair):
    eng = eng_pair
    rng = np.random.default_rng(3)
    C, G, T = 700, 20, 400
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(C, 6))
    eng.set_features(W6_NAMES)
    n = C * G
    init = np.zeros((n, 20), dtype=np.uint16)
    init[::3, 19] = 0x1EF; init[::5, 18] = 0x3F0; init[::5, 19] |= 0x3F0; init[::11, 15] = 0x201; init[::11, 16] = 0x3FE; init[::11, 17] = 0x1FF
    lines0 = (np.arange(n, dtype=np.uint32) * 13) % 57
    for pressure in (None, {"avg_lat": 120.0, "resp_time": 40.0, "move_eff": 1.05, "two_but_rot": True}):
        kw = dict(T=T, seed=5, init_rows=init, init_lines=lines0, max_moves=150, want_trace=True, want_final=True, want_scores=True, points=(3, 40, 100, 300, 1200))
        if pressure is not None:
            kw["pressure"] = pressure
        a = eng.play_games(w, G, **kw)
        assert eng.last_timing()["kernels_launched"] == 2 and eng.last_timing()["lanes"] == 2, eng.last_timing()
        b = eng.play_games(w, G, lanes=32, **kw)
        assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all()
        assert a["trace"].tobytes() == b["trace"].tobytes() and a["scores"].tobytes() == b["scores"].tobytes()


# ---------------------------------------------------------------------------------------------------------------------
# run-time specialised kernels for arbitrary feature lists (NVRTC, csrc/jit.inc)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng_jit():
    e = _engine_with_env(TB200_JIT="1")
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng_nojit():
    e = _engine_with_env(TB200_JIT="0")
    yield e
    e.close()


@pytest.mark.parametrize("kind", ["8", "holes", "floats", "one"])
def test_jit_specialised_kernels_match_oracle_and_generic(eng_jit, eng_nojit, kind):
    names = {"8": W6_NAMES + ["max_ht", "jaggedness"], "holes": ["max_ht", "row_trans", "pits", "landing_height"],
             "floats": ["mean_ht", "max_ht_diff", "min_ht_diff", "mean_pit_depth", "d_mean_ht", "pits", "col_trans", "pattern_div", "tetris_progress", "cd_4", "cuml_eroded", "deep_wells"],
             "one": ["pits"]}[kind]
    base = {"landing_height": -4.5, "eroded_cells": 3.4, "row_trans": -3.2, "col_trans": -9.3, "pits": -7.9, "cuml_wells": -3.4, "max_ht": -1.0, "jaggedness": -0.3}
    rng = np.random.default_rng(len(names))
    fids = [pyoracle.FEATURE_ID[x] for x in names]
    for C, G, T, lanes in ((2100, 6, 140, 0), (90, 4, 200, 8), (70, 3, 150, 16), (40, 2, 150, 32)):
        w = np.array([[base.get(n, 0.0) for n in names]]) + rng.normal(0, 1.5, size=(C, len(names)))
        pieces = np.array([pyoracle.piece_sequence(17 * g + 3, T) for g in range(G)], dtype=np.uint8)
        out = {}
        for tag, eng in (("jit", eng_jit), ("generic", eng_nojit)):
            eng.set_features(names)
            out[tag] = eng.play_games(w, G, pieces=pieces, lanes=lanes, want_trace=True, want_scores=True, want_final=True)
            assert eng.last_timing()["specialised"] == (1 if tag == "jit" else 0), (tag, kind, lanes)
        a, b = out["jit"], out["generic"]
        assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all() and (a["fitness"] == b["fitness"]).all()
        assert a["trace"].tobytes() == b["trace"].tobytes() and a["scores"].tobytes() == b["scores"].tobytes()
        n_o = min(C, 200)
        o = coracle.play_games(fids, w[:n_o], G, pieces=pieces, threads=THREADS, want_trace=True)
        res = a["results"][: n_o * G]
        same_results(res, o)
        for g in range(0, n_o * G, 7):
            k = int(res["pieces"][g])
            assert a["trace"][g, :k].tobytes() == o["trace"][g, :k].tobytes(), (kind, lanes, g)
    # switching lists on the same ctx: the module follows the list; w6 never uses it
    eng_jit.set_features(W6_NAMES)
    eng_jit.play_games(np.array([W6_VALS] * 64), 1, T=50, seed=1)
    assert eng_jit.last_timing()["specialised"] == 0


def test_jit_kernels_inside_the_evaluator_graph_replay(eng_jit):
    """PopulationEvaluator with a generic feature list: pinned buffers -> the repeated host-mode call is replayed as a CUDA graph whose
    play node is a run-time specialised kernel (launched through the driver API).  Three generations against the C oracle."""
    from tetris_ai_b200.learning import PopulationEvaluator
    names = W6_NAMES + ["max_ht", "jaggedness"]
    fids = [pyoracle.FEATURE_ID[x] for x in names]
    G, T, C = 5, 120, 300
    pieces = np.array([pyoracle.piece_sequence(31 * g + 7, T) for g in range(G)], dtype=np.uint8)
    ev = PopulationEvaluator(eng_jit, names, G, pieces=pieces)
    rng = np.random.default_rng(4)
    base = np.array(W6_VALS + [-0.5, -0.3])
    for gen in range(3):
        w = base[None, :] + rng.normal(0, 1.0, size=(C, len(names)))
        fit = ev.evaluate(w)
        assert eng_jit.last_timing()["specialised"] == 1
        o = coracle.play_games(fids, w, G, pieces=pieces, threads=THREADS)
        assert (fit == o["lines"].reshape(C, G).sum(axis=1) / G).all(), gen
        same_results(ev.last["results"], o)


@pytest.mark.parametrize("C,G,T,spread,names", [(3000, 5, 200, 1.0, "w6"), (13000, 4, 150, 6.0, "w6"), (500, 30, 600, 0.7, "w6"), (700, 20, 1000, 0.5, "w6")])
def test_pressure_filter_inside_the_lockstep_kernels(eng, C, G, T, spread, names):
    """move_with_pressure(move_drop) (simulator.py:88-120) for a CE population: the per-move time predicate runs inside the lockstep
    kernels (window evaluation and fixed widths), so a pressure population takes the same fast path as plain hard drops.  Against
    the per-game extended generic kernel (lanes = 8) and the C oracle; default and custom pressure parameters."""
    feat, base = W6_NAMES, np.array(W6_VALS)
    rng = np.random.default_rng(C + T)
    w = base[None, :] + rng.normal(0, spread, size=(C, len(feat)))
    fids = [pyoracle.FEATURE_ID[x] for x in feat]
    eng.set_features(feat)
    pieces = np.array([pyoracle.piece_sequence(5 * g + 11, T) for g in range(G)], dtype=np.uint8)
    for pressure in ({}, {"avg_lat": 120.0, "resp_time": 40.0, "move_eff": 1.05, "two_but_rot": True}):
        a = eng.play_games(w, G, pieces=pieces, pressure=pressure, want_trace=True, want_scores=True, want_final=True)
        assert eng.last_timing()["kernels_launched"] == 2 and eng.last_timing()["lanes"] in (1, 2, 4, 8), eng.last_timing()      # lockstep launch + fitness
        b = eng.play_games(w, G, pieces=pieces, pressure=pressure, lanes=8, want_trace=True, want_scores=True, want_final=True)
        assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all() and (a["fitness"] == b["fitness"]).all()
        assert a["trace"].tobytes() == b["trace"].tobytes() and a["scores"].tobytes() == b["scores"].tobytes()
        n_o = min(C, 250)
        o = coracle.play_games(fids, w[:n_o], G, pieces=pieces, threads=THREADS, want_trace=True, pressure=pressure)
        res = a["results"][: n_o * G]
        same_results(res, o)
        for g in range(0, n_o * G, 9):
            k = int(res["pieces"][g])
            assert a["trace"][g, :k].tobytes() == o["trace"][g, :k].tobytes(), g
    # the filter changes the play (and ends long games early: the level rises with the lines cleared)
    free = eng.play_games(w, G, pieces=pieces)["results"]
    assert int(a["results"]["pieces"].sum()) <= int(free["pieces"].sum()) and (a["results"] != free).any()
