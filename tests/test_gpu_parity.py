"""Parity of the CUDA path (through the C ABI) against the golden fixtures generated from the
unmodified reference and against the plain-C oracle on seeded inputs.  Bit-exact everywhere:
chosen placements, rows cleared, game lengths, scores (fp64 bits), all 45 features."""
import random

import numpy as np
import pytest

from oracle import coracle, pyoracle

pytestmark = pytest.mark.gpu

NAMES = list(pyoracle.FEATURE_NAMES)
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]


@pytest.fixture(scope="module")
def eng():
    from tetris_ai_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def fval(v):
    return float.fromhex(v) if isinstance(v, str) else float(v)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_same_doubles(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    ok = bits(got) == bits(want)              # strict: the sign of a zero counts too (the scorer never yields -0.0, tetris_core.cuh score_key)
    assert ok.all(), (what, np.argwhere(~ok)[:5], got[~ok][:5], want[~ok][:5])


def test_library_identifies_b200(eng):
    assert eng.sms > 0 and "NVIDIA" in eng.device_name


def test_k3_features_golden_states(eng, golden_states):
    sts = [s for s in golden_states["states"] if s["n_legal"] > 0]
    prev = np.array([s["prev"] for s in sts], dtype=np.uint16)
    r = eng.eval_features(prev, [s["piece"] for s in sts], [s["rot"] for s in sts], [s["col"] for s in sts], want_rows=True)
    assert (r["legal"] == 1).all()
    assert (r["row"] == np.array([s["row"] for s in sts])).all()
    assert (r["k"] == np.array([s["k"] for s in sts])).all()
    assert (r["rows"] == np.array([s["post"] for s in sts], dtype=np.uint16)).all()
    want = np.array([[fval(v) for v in s["features"]] for s in sts])
    assert_same_doubles(r["features"], want, "features")


def test_k3_legality_every_slot(eng, golden_states):
    nrot = [2, 4, 4, 4, 1, 2, 2]
    boards, piece, rot, col, want = [], [], [], [], []
    for st in golden_states["states"]:
        if st.get("legal_slots") is None:
            continue
        legal = {(r, c): row for r, c, row in st["legal_slots"]}
        for ro in range(nrot[st["piece"]]):
            for co in range(10):
                boards.append(st["prev"]); piece.append(st["piece"]); rot.append(ro); col.append(co)
                want.append(legal.get((ro, co), -1))
    r = eng.eval_features(np.array(boards, dtype=np.uint16), piece, rot, col)
    want = np.array(want)
    assert ((r["legal"] == 1) == (want >= 0)).all()
    assert (r["row"][want >= 0] == want[want >= 0]).all()


def test_k3_vs_coracle_random_boards(eng):
    rng = random.Random(11)
    boards, piece, rot, col = [], [], [], []
    nrot = [2, 4, 4, 4, 1, 2, 2]
    for _ in range(3000):
        hmax = rng.randint(0, 19)
        rows = [0] * 20
        for r in range(20 - hmax, 20):
            v = rng.getrandbits(10) | rng.getrandbits(10)
            rows[r] = v if v != 0x3FF else 0x3FE          # no pre-existing full rows
        p = rng.randrange(7)
        boards.append(rows); piece.append(p); rot.append(rng.randrange(nrot[p])); col.append(rng.randrange(10))
    r = eng.eval_features(np.array(boards, dtype=np.uint16), piece, rot, col, want_rows=True)
    nlegal = 0
    for i in range(len(boards)):
        o = coracle.eval_placement(boards[i], piece[i], rot[i], col[i])
        if o is None:
            assert r["legal"][i] != 1
            continue
        nlegal += 1
        assert r["legal"][i] == 1 and r["row"][i] == o["row"] and r["k"][i] == o["k"]
        assert (r["rows"][i] == o["post"]).all()
        assert_same_doubles(r["features"][i], o["features"], "board %d" % i)
    assert nlegal > 1000


def test_k1_best_move_golden_steps(eng, golden_steps):
    steps = golden_steps["steps"]
    groups = {}
    for s in steps:
        groups.setdefault(tuple(n for n, _ in s["weights"]), []).append(s)
    for names, ss in groups.items():
        eng.set_features(names)
        boards = np.array([s["board"] for s in ss], dtype=np.uint16)
        piece = [s["piece"] for s in ss]
        nxt = [s["next"] for s in ss]
        w = np.array([[x for _, x in s["weights"]] for s in ss])
        for key, la in (("l1", 1), ("l2", 2)):
            r = eng.best_move(boards, piece, nxt if la == 2 else None, w, lookahead=la, want_rows=True)
            for i, s in enumerate(ss):
                want = s[key]
                if want is None:
                    if key == "l1":
                        assert r["move"][i][1] == 255 and r["count"][i] == 0
                    continue
                assert tuple(int(x) for x in r["move"][i]) == (want["rot"], want["row"], want["col"], want["k"]), (key, i)
                assert_same_doubles([r["score"][i]], [float.fromhex(want["score"])], "score")
                if la == 1:
                    assert r["count"][i] == want["n_legal"]
                elif want["n_leaves"]:
                    assert r["count"][i] == want["n_leaves"]


@pytest.mark.parametrize("ms,bit", [("drop", 0), ("slide", 1), ("wiggle_top", 4), ("wiggle_drop", 8)])
@pytest.mark.parametrize("lookahead", [1, 2])
def test_k1_best_move_ex_move_sets_vs_coracle(eng, golden_steps, ms, bit, lookahead):
    """tb200_best_move_ex: one decision per board under every move generator, with and without the pressure filter and
    with lines already cleared (the level the filter reads) — against the C oracle's one-move game on the same board."""
    names = ["max_ht", "row_trans", "pits", "landing_height", "cuml_wells", "col_trans"]
    fids = [pyoracle.FEATURE_ID[n] for n in names]
    rng = random.Random(17 + bit + lookahead)
    boards = [s["board"] for s in golden_steps["steps"]][:60]
    for _ in range(40):                                   # cave-rich random boards: slides and wiggles find extra moves
        hmax = rng.randint(2, 14)
        rows = [0] * 20
        for r in range(20 - hmax, 20):
            v = rng.getrandbits(10) | rng.getrandbits(10)
            rows[r] = v if v != 0x3FF else 0x37F
        boards.append(rows)
    N = len(boards)
    boards = np.array(boards, dtype=np.uint16)
    piece = [rng.randrange(7) for _ in range(N)]
    nxt = [rng.randrange(7) for _ in range(N)]
    w = np.array([[rng.gauss(0, 2.0) - 1.0 for _ in names] for _ in range(N)])
    lines = np.array([rng.choice([0, 0, 7, 35, 60, 95]) for _ in range(N)], dtype=np.uint32)
    eng.set_features(names)
    for pressure in (None, {"avg_lat": 200.0, "resp_time": 60.0, "move_eff": 1.1, "two_but_rot": True}):
        r = eng.best_move(boards, piece, nxt if lookahead == 2 else None, w, lookahead=lookahead, want_rows=True,
                          move_set=bit, pressure=pressure, lines=lines)
        for i in range(N):
            seq = [piece[i], nxt[i]] if lookahead == 2 else [piece[i]]
            o = coracle.play_one(fids, w[i], seq, lookahead, init_rows=boards[i], move_set=ms, pressure=pressure, init_lines=int(lines[i]))
            if len(o["trace"]) == 0:
                assert r["move"][i][1] == 255, (i, pressure)
                continue
            assert tuple(int(x) for x in r["move"][i]) == tuple(int(x) for x in o["trace"][0]), (i, pressure)


@pytest.mark.parametrize("lanes", [64, 32, 16, 8])
def test_k2_golden_games(eng, golden_games, lanes):
    for g in golden_games["games"]:
        if g.get("tie") == "last":
            continue
        eng.set_features([n for n, _ in g["weights"]])
        w = np.array([[x for _, x in g["weights"]]])
        seq = np.array([pyoracle.piece_sequence(g["seed"], g["T"])], dtype=np.uint8)
        r = eng.play_games(w, 1, pieces=seq, lookahead=g["lookahead"], lanes=lanes, want_trace=True, want_final=True)
        res = r["results"][0]
        assert int(res["pieces"]) == g["pieces"] and int(res["lines"]) == g["lines"]
        assert [int(x) for x in res["hist"]] == g["hist"]
        assert int(res["score"]) == g["score"] and int(res["placements"]) == g["scorer_calls"]
        assert r["trace"][0, :g["pieces"]].tobytes().hex() == g["trace_hex"]
        assert list(r["final"][0]) == g["final_board"]
        assert r["fitness"][0] == g["lines"] / 1


def _weight_sets(rng, n, names, kind):
    if kind == "w6":
        return [[rng.gauss(v, 1.0) for v in W6_VALS] for _ in range(n)]
    if kind == "wide":
        return [[rng.gauss(0, 10.0) for _ in names] for _ in range(n)]
    return [[(W6_VALS[W6_NAMES.index(x)] if x in W6_NAMES else 0.0) + rng.gauss(0, 0.3) for x in names] for _ in range(n)]


@pytest.mark.parametrize("lanes", [64, 32, 16, 8])
@pytest.mark.parametrize("names,kind,lookahead,C,G,T", [
    (W6_NAMES, "w6", 1, 48, 4, 300),
    (W6_NAMES, "wide", 1, 96, 4, 300),
    (NAMES, "all45", 1, 24, 3, 200),
    (["pits", "max_ht", "d_mean_ht", "mean_pit_depth", "pattern_div", "tetris_progress", "cd_4"], "wide", 1, 40, 3, 200),
    (W6_NAMES, "w6", 2, 6, 2, 40),
    (NAMES, "all45", 2, 4, 2, 25),
])
def test_k2_vs_coracle(eng, lanes, names, kind, lookahead, C, G, T):
    rng = random.Random(hash((kind, lookahead, C)) & 0xFFFF)
    fids = [pyoracle.FEATURE_ID[n] for n in names]
    w = np.array(_weight_sets(rng, C, names, kind))
    eng.set_features(names)
    seed = 4242 + lookahead
    r = eng.play_games(w, G, T=T, seed=seed, lookahead=lookahead, lanes=lanes, want_trace=True)
    o = coracle.play_games(fids, w, G, T=T, seed=seed, lookahead=lookahead, threads=16, want_trace=True)
    res = r["results"]
    assert (res["pieces"] == o["pieces"]).all()
    assert (res["lines"] == o["lines"]).all()
    assert (res["hist"] == o["hist"]).all()
    assert (res["score"] == o["score"]).all()
    assert (res["placements"] == o["placements"]).all()
    for g in range(C * G):
        n = int(res["pieces"][g])
        assert (r["trace"][g, :n] == o["trace"][g, :n]).all(), g
    fit = o["lines"].reshape(C, G).sum(axis=1) / G
    assert (r["fitness"] == fit).all()


def test_k2_ce_generation_golden(eng, golden_ce):
    """BASELINE config 2: generation 0 of the golden CE run (100 candidates x 10 games, T=1000)."""
    ce = golden_ce["ce"]
    rng = random.Random(ce["rng_seed"])
    F = len(ce["features"])
    pop = [[rng.gauss(0, ce["stdev"]) for _ in range(F)] for _ in range(ce["width"])]
    seqs = np.array([pyoracle.piece_sequence(ce["seed0"] + g, ce["T"]) for g in range(ce["games"])], dtype=np.uint8)
    eng.set_features(ce["features"])
    r = eng.play_games(np.array(pop), ce["games"], pieces=seqs)
    want = [float.fromhex(x) for x in ce["generations"][0]["fitness"]]
    assert list(r["fitness"]) == want
    assert int(r["results"]["placements"].sum()) == 775946 and int(r["results"]["pieces"].sum()) == 40087


def test_k4_hashed_pieces(eng):
    got = eng.hashed_pieces(seed=99, first_stream=5, n_streams=7, T=300)
    for s in range(7):
        assert list(got[s]) == [pyoracle.hashed_piece(99, 5 + s, t) for t in range(300)]


def test_k2_many_games_vs_coracle(eng):
    """20 000 concurrent games, each with its own candidate and its own hashed stream (config 5 shape)."""
    rng = np.random.default_rng(99)
    C = 20000
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(C, 6))
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    eng.set_features(W6_NAMES)
    seqid = np.arange(C, dtype=np.int32)
    r = eng.play_games(w, 1, T=150, seed=7, seq_of_game=seqid)
    o = coracle.play_games(fids, w, 1, T=150, seed=7, seq_of_game=seqid, threads=16)
    res = r["results"]
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all()
    assert (res["score"] == o["score"]).all() and (res["placements"] == o["placements"]).all()


def test_properties_full_size(eng):
    """Size-independent properties at a size the oracle would not finish quickly:
    lanes 8 / 16 / 32 give identical results; results do not depend on how games are batched."""
    rng = np.random.default_rng(5)
    C, G, T = 2000, 8, 600
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 2.0, size=(C, 6))
    eng.set_features(W6_NAMES)
    base = eng.play_games(w, G, T=T, seed=11, lanes=32)["results"]
    for lanes in (64, 16, 8):
        other = eng.play_games(w, G, T=T, seed=11, lanes=lanes)["results"]
        assert (other == base).all()
    half = eng.play_games(w[:C // 2], G, T=T, seed=11)["results"]
    assert (half == base[: C // 2 * G]).all()
    assert (base["lines"] == (base["hist"] * np.array([1, 2, 3, 4])).sum(axis=1)).all()
    assert (base["pieces"] <= T).all()


@pytest.mark.parametrize("n,T,spread", [(1000, 400, 6.0), (600, 300, 8.0), (2300, 200, 5.0), (400, 256, 10.0), (1000, 1000, 0.5)])
def test_k2_respread_stages_vs_single_launch_and_coracle(eng, n, T, spread):
    """Latency regime (automatic lanes, 2..16 games per SM): the play is split into stream-ordered stages that hand the
    surviving games — board, counters, position in the piece stream — to sparser launches.  Every output must equal the
    single-launch kernel's (lanes = 32) and the C oracle's: results, every move, every winning score, final boards."""
    rng = np.random.default_rng(n + T)
    w = np.array(W6_VALS)[None, :] + rng.normal(0, spread, size=(n, 6))          # wide spread: games die at very different ages
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    eng.set_features(W6_NAMES)
    seqid = (np.arange(n, dtype=np.int32) * 7) % 64
    init = np.zeros((n, 20), dtype=np.uint16)
    init[::3, 19] = 0x1EF; init[::5, 18] = 0x3F0; init[::5, 19] |= 0x3F0
    lines0 = (np.arange(n, dtype=np.uint32) * 13) % 57
    kw = dict(T=T, seed=21, seq_of_game=seqid, init_rows=init, init_lines=lines0, want_trace=True, want_scores=True, want_final=True)
    a = eng.play_games(w, 1, **kw)                       # staged
    assert eng.last_timing()["kernels_launched"] >= 3
    b = eng.play_games(w, 1, lanes=32, **kw)             # one launch
    assert eng.last_timing()["kernels_launched"] <= 2
    assert (a["results"] == b["results"]).all()
    assert (a["final"] == b["final"]).all()
    for g in range(n):
        k = int(a["results"]["pieces"][g])
        assert a["trace"][g, :k].tobytes() == b["trace"][g, :k].tobytes()
        assert bits(a["scores"][g, :k]).tobytes() == bits(b["scores"][g, :k]).tobytes()
    # the oracle's batch entry starts from empty boards: same population without the initial boards / lines
    res = eng.play_games(w, 1, T=T, seed=21, seq_of_game=seqid)["results"]
    assert eng.last_timing()["kernels_launched"] >= 3
    o = coracle.play_games(fids, w, 1, T=T, seed=21, seq_of_game=seqid, threads=16)
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all()
    assert (res["score"] == o["score"]).all() and (res["placements"] == o["placements"]).all()
    assert (res["hist"] == o["hist"]).all()


@pytest.mark.parametrize("C,G,T,spread,explicit", [(500, 30, 200, 2.0, True), (3001, 5, 131, 6.0, True), (2000, 7, 160, 3.0, False), (13000, 4, 128, 8.0, True)])
def test_k2_lockstep_population_vs_coracle(eng, C, G, T, spread, explicit):
    """Population regime (more than 80 games per SM on G shared sequences, automatic lanes): the lockstep kernel K2s plays
    stage by stage and re-packs the survivors of each sequence into full warps.  Ragged sizes (C not a multiple of the games
    per warp, T not a multiple of the stage length), explicit and counter-based sequences: results, moves and winning
    scores must equal the per-game kernel's (lanes = 8) and the C oracle's."""
    rng = np.random.default_rng(C + G)
    w = np.array(W6_VALS)[None, :] + rng.normal(0, spread, size=(C, 6))
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    eng.set_features(W6_NAMES)
    pieces = None
    if explicit:
        pieces = np.array([pyoracle.piece_sequence(7 * g + 1, T) for g in range(G)], dtype=np.uint8)
    kw = dict(pieces=pieces, T=T, seed=77)
    a = eng.play_games(w, G, want_trace=True, want_scores=True, want_final=True, **kw)
    assert eng.last_timing()["kernels_launched"] == 2 and eng.last_timing()["lanes"] in (1, 2, 4, 8)
    b = eng.play_games(w, G, lanes=8, want_trace=True, want_scores=True, want_final=True, **kw)
    assert (a["results"] == b["results"]).all()
    assert (a["final"] == b["final"]).all()
    assert (a["fitness"] == b["fitness"]).all()
    for g in range(0, C * G, 7):
        k = int(a["results"]["pieces"][g])
        assert a["trace"][g, :k].tobytes() == b["trace"][g, :k].tobytes()
        assert bits(a["scores"][g, :k]).tobytes() == bits(b["scores"][g, :k]).tobytes()
    o = coracle.play_games(fids, w, G, pieces=pieces, T=T, seed=77, threads=16, want_trace=True)
    res = a["results"]
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all()
    assert (res["score"] == o["score"]).all() and (res["placements"] == o["placements"]).all()
    assert (res["hist"] == o["hist"]).all()
    for g in range(0, C * G, 11):
        k = int(res["pieces"][g])
        assert a["trace"][g, :k].tobytes() == o["trace"][g, :k].tobytes()


def test_k2_lockstep_one_lane_per_game_full_population(eng):
    """The whole BASELINE configs[2] population (10 000 candidates x 30 games) on one GPU: K2s runs with ONE lane per game
    (warp-uniform descriptors).  Every result against the per-game kernel; the first 150 candidates against the C oracle."""
    rng = np.random.default_rng(10)
    C, G, T = 10000, 30, 130
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 2.5, size=(C, 6))
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    eng.set_features(W6_NAMES)
    pieces = np.array([pyoracle.piece_sequence(13 * g + 5, T) for g in range(G)], dtype=np.uint8)
    a = eng.play_games(w, G, pieces=pieces, want_final=True)
    assert eng.last_timing()["kernels_launched"] == 2 and eng.last_timing()["lanes"] == 1
    b = eng.play_games(w, G, pieces=pieces, lanes=8, want_final=True)
    assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all() and (a["fitness"] == b["fitness"]).all()
    o = coracle.play_games(fids, w[:150], G, pieces=pieces, threads=16)
    res = a["results"][:150 * G]
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all()
    assert (res["score"] == o["score"]).all() and (res["placements"] == o["placements"]).all() and (res["hist"] == o["hist"]).all()


@pytest.mark.parametrize("kind,C,G,T", [("8", 2100, 6, 140), ("all45", 1600, 8, 130), ("holes", 13000, 30, 128)])
def test_k2_lockstep_other_feature_lists_vs_coracle(eng, kind, C, G, T):
    """The lockstep kernel with feature lists other than w6 (general placement / feature / scorer code, run-time list or all 45):
    against the per-game generic kernel (lanes = 8) and the C oracle."""
    names = {"8": W6_NAMES + ["max_ht", "jaggedness"], "all45": NAMES, "holes": ["max_ht", "row_trans", "pits", "landing_height"]}[kind]
    rng = np.random.default_rng(len(names) + C)
    base = {"landing_height": -4.5, "eroded_cells": 3.4, "row_trans": -3.2, "col_trans": -9.3, "pits": -7.9, "cuml_wells": -3.4, "max_ht": -1.0, "jaggedness": -0.3}
    w = np.array([[base.get(n, 0.0) for n in names]]) + rng.normal(0, 1.5 if kind != "all45" else 0.4, size=(C, len(names)))
    fids = [pyoracle.FEATURE_ID[x] for x in names]
    eng.set_features(names)
    pieces = np.array([pyoracle.piece_sequence(17 * g + 3, T) for g in range(G)], dtype=np.uint8)
    a = eng.play_games(w, G, pieces=pieces, want_trace=True, want_scores=True, want_final=True)
    assert eng.last_timing()["kernels_launched"] == 2 and eng.last_timing()["lanes"] in (1, 2, 4, 8)
    b = eng.play_games(w, G, pieces=pieces, lanes=8, want_trace=True, want_scores=True, want_final=True)
    assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all() and (a["fitness"] == b["fitness"]).all()
    for g in range(0, C * G, 53):
        k = int(a["results"]["pieces"][g])
        assert a["trace"][g, :k].tobytes() == b["trace"][g, :k].tobytes()
        assert bits(a["scores"][g, :k]).tobytes() == bits(b["scores"][g, :k]).tobytes()
    n_o = min(C, 300)
    o = coracle.play_games(fids, w[:n_o], G, pieces=pieces, threads=16)
    res = a["results"][:n_o * G]
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all()
    assert (res["score"] == o["score"]).all() and (res["placements"] == o["placements"]).all() and (res["hist"] == o["hist"]).all()


def test_k2_lockstep_initial_boards_lines_and_max_moves(eng):
    """K2s with initial boards, lines already cleared and a max_moves stop that is not a stage boundary."""
    rng = np.random.default_rng(3)
    C, G, T = 700, 20, 400
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(C, 6))
    eng.set_features(W6_NAMES)
    n = C * G
    init = np.zeros((n, 20), dtype=np.uint16)
    init[::3, 19] = 0x1EF; init[::5, 18] = 0x3F0; init[::5, 19] |= 0x3F0
    lines0 = (np.arange(n, dtype=np.uint32) * 13) % 57
    kw = dict(T=T, seed=5, init_rows=init, init_lines=lines0, max_moves=150, want_trace=True, want_final=True)
    a = eng.play_games(w, G, **kw)
    assert eng.last_timing()["kernels_launched"] == 2
    b = eng.play_games(w, G, lanes=32, **kw)
    assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all()
    assert (a["results"]["pieces"] <= 150).all()
    for g in range(0, n, 5):
        k = int(a["results"]["pieces"][g])
        assert a["trace"][g, :k].tobytes() == b["trace"][g, :k].tobytes()
    # explicit sequences, two of which end early (a piece id > 6 ends the game there — an extension of this library, so
    # the comparison is against the per-game kernel only)
    pieces = np.array([pyoracle.piece_sequence(9 * g + 2, T) for g in range(G)], dtype=np.uint8)
    pieces[3, 200:] = 7
    pieces[G - 1, 64:] = 7
    a = eng.play_games(w, G, pieces=pieces, want_final=True)
    assert eng.last_timing()["kernels_launched"] == 2
    b = eng.play_games(w, G, pieces=pieces, lanes=16, want_final=True)
    assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all()
    assert (a["results"]["pieces"].reshape(C, G)[:, 3] <= 200).all() and (a["results"]["pieces"].reshape(C, G)[:, G - 1] <= 64).all()


@pytest.mark.parametrize("n,T,explicit", [(33000, 70, False), (70001, 40, True), (300000, 24, False)])
def test_k2_piece_binned_kernel_vs_coracle(eng, n, T, explicit):
    """Independent piece streams, >= 32 768 games, automatic lanes: the piece-binned kernel K2b moves every game through
    device-wide per-piece-type queues, one piece per task.  Results, moves, winning scores and final boards must equal the
    per-game kernel's (lanes = 4) and the C oracle's; with explicit sequences some of them end early."""
    rng = np.random.default_rng(n)
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 4.0, size=(n, 6))
    fids = [pyoracle.FEATURE_ID[x] for x in W6_NAMES]
    eng.set_features(W6_NAMES)
    kw = dict(T=T, seed=31)
    if explicit:
        S = 257
        pieces = np.array([[random.Random(3 * g + 1).randrange(7) for _ in range(T)] for g in range(S)], dtype=np.uint8)
        kw = dict(pieces=pieces, seq_of_game=((np.arange(n, dtype=np.int64) * 11) % S).astype(np.int32))
    else:
        kw["seq_of_game"] = np.arange(n, dtype=np.int32)
    a = eng.play_games(w, 1, want_trace=True, want_scores=True, want_final=True, **kw)
    assert eng.last_timing()["kernels_launched"] == 3 and eng.last_timing()["lanes"] in (1, 2, 4)       # init + K2b + fitness
    b = eng.play_games(w, 1, lanes=4, want_trace=True, want_scores=True, want_final=True, **kw)
    assert eng.last_timing()["kernels_launched"] == 2
    assert (a["results"] == b["results"]).all()
    assert (a["final"] == b["final"]).all()
    assert (a["fitness"] == b["fitness"]).all()
    for g in range(0, n, 97):
        k = int(a["results"]["pieces"][g])
        assert a["trace"][g, :k].tobytes() == b["trace"][g, :k].tobytes()
        assert bits(a["scores"][g, :k]).tobytes() == bits(b["scores"][g, :k]).tobytes()
    o = coracle.play_games(fids, w, 1, threads=16, want_trace=True, **{k2: v for k2, v in kw.items()})
    res = a["results"]
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all()
    assert (res["score"] == o["score"]).all() and (res["placements"] == o["placements"]).all()
    assert (res["hist"] == o["hist"]).all()
    for g in range(0, n, 89):
        k = int(res["pieces"][g])
        assert a["trace"][g, :k].tobytes() == o["trace"][g, :k].tobytes()


def test_k2_piece_binned_initial_boards_lines_max_moves_and_short_sequences(eng):
    rng = np.random.default_rng(8)
    n, T = 40000, 90
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 3.0, size=(n, 6))
    eng.set_features(W6_NAMES)
    init = np.zeros((n, 20), dtype=np.uint16)
    init[::3, 19] = 0x1EF; init[::5, 18] = 0x3F0; init[::5, 19] |= 0x3F0
    init[::7, 2:] = 0x2AA                                            # tall boards: some games end at once
    lines0 = (np.arange(n, dtype=np.uint32) * 13) % 57
    S = 64
    pieces = np.array([[random.Random(5 * g + 2).randrange(7) for _ in range(T)] for g in range(S)], dtype=np.uint8)
    pieces[5, 30:] = 7; pieces[6, 0:] = 7                            # a sequence that ends early and an empty one
    kw = dict(pieces=pieces, seq_of_game=(np.arange(n, dtype=np.int32) % S), init_rows=init, init_lines=lines0, max_moves=50,
              want_trace=True, want_final=True)
    a = eng.play_games(w, 1, **kw)
    assert eng.last_timing()["kernels_launched"] == 3
    b = eng.play_games(w, 1, lanes=32, **kw)
    assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all()
    assert (a["results"]["pieces"] <= 50).all() and (a["results"]["pieces"][6::S] == 0).all()
    assert (a["results"]["lines"][6::S] == lines0[6::S]).all() and (a["final"][6::S] == init[6::S]).all()    # nothing played: state untouched
    for lanes in (8, 2):                                              # the grouped kernel reports untouched games the same way
        b2 = eng.play_games(w, 1, lanes=lanes, **kw)
        assert (a["results"] == b2["results"]).all() and (a["final"] == b2["final"]).all()
    for g in range(0, n, 31):
        k = int(a["results"]["pieces"][g])
        assert a["trace"][g, :k].tobytes() == b["trace"][g, :k].tobytes()


def test_k2_lockstep_and_binned_edge_shapes(eng):
    """Corner cases of the two queue-driven kernels against the per-game kernel: the smallest G and stage count K2s takes,
    a population in which most games die within the first stage (whole buckets stay empty), K2b at its smallest size with
    one-piece and two-piece games, and K2b on boards where nothing can be placed."""
    eng.set_features(W6_NAMES)
    rng = np.random.default_rng(21)

    def same(kw_a, kw_b, *args, **kw):
        a = eng.play_games(*args, want_final=True, **kw, **kw_a)
        la = eng.last_timing()["kernels_launched"]
        b = eng.play_games(*args, want_final=True, **kw, **kw_b)
        assert (a["results"] == b["results"]).all() and (a["final"] == b["final"]).all() and (a["fitness"] == b["fitness"]).all()
        return a, la

    # K2s: G = 4, exactly two full stages' worth of pieces
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 2.0, size=(3100, 6))
    _, la = same({}, {"lanes": 8}, w, 4, T=128, seed=3)
    assert la == 2
    # K2s: hopeless candidates (they stack up and die in a few dozen pieces), a few good ones
    w = rng.normal(0, 10.0, size=(2500, 6)); w[::97] = np.array(W6_VALS)
    a, la = same({}, {"lanes": 16}, w, 6, T=300, seed=4)
    assert la == 2 and a["results"]["pieces"].min() < 40 and a["results"]["pieces"].max() == 300
    # K2b: smallest population, one and two pieces per game
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 2.0, size=(32768, 6))
    ids = np.arange(32768, dtype=np.int32)
    for T in (1, 2):
        _, la = same({}, {"lanes": 8}, w, 1, T=T, seed=6, seq_of_game=ids)
        assert la == 3
    # K2b: boards filled to the top row (checkerboard, nothing ever clears): most pieces do not fit at all
    init = np.zeros((32768, 20), dtype=np.uint16); init[:, 1::2] = 0x2AA; init[:, 2::2] = 0x155
    a, la = same({}, {"lanes": 32}, w, 1, T=30, seed=7, seq_of_game=ids, init_rows=init)
    assert la == 3 and (a["results"]["pieces"] < 30).all() and (a["results"]["pieces"] == 0).any()


# ------------------------------------------------------------------------------------------------
# edge cases: tiny / ragged shapes, immediate game over, max_moves, explicit sequence maps, many tickets
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lanes", [0, 64, 32, 16, 8])
def test_edge_shapes(eng, lanes):
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    eng.set_features(W6_NAMES)
    # one candidate, one game, one piece
    r = eng.play_games(np.array([W6_VALS]), 1, pieces=np.array([[3]], dtype=np.uint8), lanes=lanes, want_trace=True)
    o = coracle.play_one(fids, W6_VALS, [3], 1)
    assert int(r["results"]["pieces"][0]) == 1 and r["trace"][0, 0].tobytes() == o["trace"][0].tobytes()
    # ragged: 7 candidates x 3 games (21 games: not a multiple of any group count), explicit sequence map, S > G
    rng = np.random.default_rng(17)
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 4.0, size=(7, 6))
    seqs = rng.integers(0, 7, size=(5, 90)).astype(np.uint8)
    smap = rng.integers(0, 5, size=21).astype(np.int32)
    r = eng.play_games(w, 3, pieces=seqs, seq_of_game=smap, lanes=lanes, want_trace=True, want_final=True)
    for g in range(21):
        o = coracle.play_one(fids, w[g // 3], seqs[smap[g]], 1)
        n = o["pieces"]
        assert int(r["results"]["pieces"][g]) == n and int(r["results"]["score"][g]) == o["score"]
        assert r["trace"][g, :n].tobytes() == o["trace"].tobytes() and list(r["final"][g]) == list(o["final"])


@pytest.mark.parametrize("lanes", [64, 32, 8])
@pytest.mark.parametrize("lookahead", [1, 2])
def test_edge_initial_boards_and_max_moves(eng, lanes, lookahead):
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    eng.set_features(W6_NAMES)
    rng = np.random.default_rng(23)
    boards = []
    tall = [0x1FF] * 19 + [0x3FE]                 # heights 19..: only a vertical I in the last column or nothing fits
    tall[0] = 0x0FF
    boards.append(tall)
    boards.append([0] * 12 + [0x155, 0x2AA, 0x155, 0x2AA, 0x3FE, 0x1FF, 0x3FD, 0x37F])
    boards.append([0] * 20)
    boards = np.array(boards, dtype=np.uint16)
    seqs = rng.integers(0, 7, size=(3, 40)).astype(np.uint8)
    w = np.array(W6_VALS)[None, :] + rng.normal(0, 1.0, size=(3, 6))
    for max_moves in (0, 1, 7):
        r = eng.play_games(w, 1, pieces=seqs, seq_of_game=np.arange(3, dtype=np.int32), init_rows=boards, lookahead=lookahead,
                           lanes=lanes, max_moves=max_moves, want_trace=True, want_final=True)
        for g in range(3):
            # the oracle has no max_moves: a game stopped after m moves equals the first m moves of the full game
            o = coracle.play_one(fids, w[g], seqs[g], lookahead, init_rows=boards[g])
            n = int(r["results"]["pieces"][g])
            if max_moves == 0:
                assert n == o["pieces"] and list(r["final"][g]) == list(o["final"]) and int(r["results"]["lines"][g]) == o["lines"]
            else:
                assert n == min(max_moves, o["pieces"])
            if lookahead == 1 or max_moves == 0:
                assert r["trace"][g, :n].tobytes() == o["trace"][:n].tobytes()


def test_edge_many_tickets_and_determinism(eng):
    """More games than resident groups for the wide kernels (persistent blocks take several tickets); repeated calls agree."""
    eng.set_features(W6_NAMES)
    rng = np.random.default_rng(31)
    C = 6000
    w = rng.normal(0, 10.0, size=(C, 6))
    a = eng.play_games(w, 1, T=120, seed=3, seq_of_game=np.arange(C, dtype=np.int32), lanes=64)["results"]
    b = eng.play_games(w, 1, T=120, seed=3, seq_of_game=np.arange(C, dtype=np.int32), lanes=32)["results"]
    c = eng.play_games(w, 1, T=120, seed=3, seq_of_game=np.arange(C, dtype=np.int32), lanes=8)["results"]
    d = eng.play_games(w, 1, T=120, seed=3, seq_of_game=np.arange(C, dtype=np.int32), lanes=64)["results"]
    assert (a == b).all() and (a == c).all() and (a == d).all()
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    o = coracle.play_games(fids, w[:500], 1, T=120, seed=3, seq_of_game=np.arange(500, dtype=np.int32), threads=16)
    assert (a["pieces"][:500] == o["pieces"]).all() and (a["score"][:500] == o["score"]).all()


def test_edge_invalid_piece_ends_game(eng):
    """A piece id outside 0..6 is treated as the end of the sequence (documented in tetris_b200.h)."""
    eng.set_features(W6_NAMES)
    seq = np.array([[1, 2, 9, 3, 4]], dtype=np.uint8)
    for lanes in (64, 32, 16, 8):
        r = eng.play_games(np.array([W6_VALS]), 1, pieces=seq, lanes=lanes)
        assert int(r["results"]["pieces"][0]) == 2


def test_k1_large_batch(eng, golden_steps):
    steps = [s for s in golden_steps["steps"] if s["l1"] is not None and [n for n, _ in s["weights"]] == W6_NAMES]
    assert len(steps) > 20
    reps = 3000 // len(steps) + 1
    ss = (steps * reps)[:3000]
    eng.set_features(W6_NAMES)
    r = eng.best_move(np.array([s["board"] for s in ss], dtype=np.uint16), [s["piece"] for s in ss], None,
                      np.array([[x for _, x in s["weights"]] for s in ss]), lookahead=1)
    for i, s in enumerate(ss):
        want = s["l1"]
        assert tuple(int(x) for x in r["move"][i]) == (want["rot"], want["row"], want["col"], want["k"])


# ------------------------------------------------------------------------------------------------
# move_with_overhang_slide(move_drop) on the device
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lanes", [32, 16, 8])
def test_slide_golden_games(eng, lanes):
    from conftest import load_golden
    for g in load_golden("games_slide.json")["games"]:
        eng.set_features([n for n, _ in g["weights"]])
        w = np.array([[x for _, x in g["weights"]]])
        seq = np.array([pyoracle.piece_sequence(g["seed"], g["T"])], dtype=np.uint8)
        r = eng.play_games(w, 1, pieces=seq, lookahead=g["lookahead"], lanes=lanes, want_trace=True, want_final=True, move_set=1)
        res = r["results"][0]
        assert int(res["pieces"]) == g["pieces"] and int(res["lines"]) == g["lines"] and [int(x) for x in res["hist"]] == g["hist"]
        assert int(res["score"]) == g["score"] and int(res["placements"]) == g["scorer_calls"]
        assert r["trace"][0, :g["pieces"]].tobytes().hex() == g["trace_hex"]
        assert list(r["final"][0]) == g["final_board"]


@pytest.mark.parametrize("lanes", [0, 32, 8])
@pytest.mark.parametrize("names,kind,lookahead,C,G,T", [
    (W6_NAMES, "wide", 1, 64, 3, 250),
    (["max_ht", "row_trans", "pits", "landing_height"], "holes", 1, 64, 3, 250),
    (NAMES, "all45", 1, 16, 2, 150),
    (["max_ht", "row_trans", "pits", "landing_height"], "holes", 2, 6, 2, 30),
])
def test_slide_vs_coracle(eng, lanes, names, kind, lookahead, C, G, T):
    rng = random.Random(hash((kind, lookahead, C, "slide")) & 0xFFFF)
    fids = [pyoracle.FEATURE_ID[n] for n in names]
    if kind == "holes":
        w = np.array([[-1.0, rng.gauss(-0.2, 0.2), rng.gauss(0.4, 0.3), -0.5] for _ in range(C)])
    else:
        w = np.array(_weight_sets(rng, C, names, kind))
    eng.set_features(names)
    r = eng.play_games(w, G, T=T, seed=77, lookahead=lookahead, lanes=lanes, want_trace=True, move_set=1)
    o = coracle.play_games(fids, w, G, T=T, seed=77, lookahead=lookahead, threads=16, want_trace=True, move_set="slide")
    d = coracle.play_games(fids, w, G, T=T, seed=77, lookahead=lookahead, threads=16, move_set="drop")
    res = r["results"]
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all() and (res["score"] == o["score"]).all()
    assert (res["placements"] == o["placements"]).all()
    for g in range(C * G):
        n = int(res["pieces"][g])
        assert (r["trace"][g, :n] == o["trace"][g, :n]).all(), g
    assert (o["placements"] != d["placements"]).any()          # slides really occurred


# ------------------------------------------------------------------------------------------------
# move_with_pressure on the device
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lanes", [32, 8])
def test_pressure_golden_games(eng, lanes):
    from conftest import load_golden
    for g in load_golden("games_pressure.json")["games"]:
        eng.set_features([n for n, _ in g["weights"]])
        w = np.array([[x for _, x in g["weights"]]])
        seq = np.array([pyoracle.piece_sequence(g["seed"], g["T"])], dtype=np.uint8)
        ms = 1 if g["move_gen"] == "overhang_slide" else 0
        r = eng.play_games(w, 1, pieces=seq, lookahead=g["lookahead"], lanes=lanes, want_trace=True, want_final=True,
                           move_set=ms, pressure=g["pressure"])
        res = r["results"][0]
        assert int(res["pieces"]) == g["pieces"] and int(res["lines"]) == g["lines"] and [int(x) for x in res["hist"]] == g["hist"]
        assert int(res["score"]) == g["score"] and int(res["placements"]) == g["scorer_calls"]
        assert r["trace"][0, :g["pieces"]].tobytes().hex() == g["trace_hex"]
        assert list(r["final"][0]) == g["final_board"]


@pytest.mark.parametrize("lookahead,C,G,T,ms", [(1, 96, 3, 500, 0), (1, 64, 3, 400, 1), (2, 6, 2, 40, 1)])
def test_pressure_vs_coracle(eng, lookahead, C, G, T, ms):
    rng = random.Random(4000 + lookahead + ms)
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    w = np.array(_weight_sets(rng, C, W6_NAMES, "w6"))
    pr = {"avg_lat": 210.0, "resp_time": 70.0, "move_eff": 1.3, "two_but_rot": True}
    eng.set_features(W6_NAMES)
    r = eng.play_games(w, G, T=T, seed=12, lookahead=lookahead, want_trace=True, move_set=ms, pressure=pr)
    o = coracle.play_games(fids, w, G, T=T, seed=12, lookahead=lookahead, threads=16, want_trace=True, move_set=ms, pressure=pr)
    res = r["results"]
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all() and (res["score"] == o["score"]).all()
    assert (res["placements"] == o["placements"]).all()
    for g in range(C * G):
        n = int(res["pieces"][g])
        assert (r["trace"][g, :n] == o["trace"][g, :n]).all(), g
    if lookahead == 1:
        assert (res["pieces"] < T).any()                    # the pressure filter ended some games


# ------------------------------------------------------------------------------------------------
# move_wiggle / move_with_wiggle(move_drop) on the device
# ------------------------------------------------------------------------------------------------
def test_wiggle_golden_games(eng):
    from conftest import load_golden
    for g in load_golden("games_wiggle.json")["games"]:
        eng.set_features([n for n, _ in g["weights"]])
        w = np.array([[x for _, x in g["weights"]]])
        seq = np.array([pyoracle.piece_sequence(g["seed"], g["T"])], dtype=np.uint8)
        ms = 4 if g["move_gen"] == "wiggle_top" else 8
        r = eng.play_games(w, 1, pieces=seq, lookahead=g["lookahead"], want_trace=True, want_final=True, move_set=ms, pressure=g["pressure"])
        res = r["results"][0]
        assert int(res["pieces"]) == g["pieces"] and int(res["lines"]) == g["lines"] and [int(x) for x in res["hist"]] == g["hist"]
        assert int(res["score"]) == g["score"] and int(res["placements"]) == g["scorer_calls"]
        assert r["trace"][0, :g["pieces"]].tobytes().hex() == g["trace_hex"]
        assert list(r["final"][0]) == g["final_board"]


@pytest.mark.parametrize("names,kind,ms,C,G,T", [
    (W6_NAMES, "wide", "wiggle_top", 64, 3, 200),
    (["max_ht", "row_trans", "pits", "landing_height"], "holes", "wiggle_top", 64, 3, 200),
    (["max_ht", "row_trans", "pits", "landing_height"], "holes", "wiggle_drop", 64, 3, 200),
    (NAMES, "all45", "wiggle_drop", 16, 2, 120),
])
def test_wiggle_vs_coracle(eng, names, kind, ms, C, G, T):
    rng = random.Random(hash((kind, ms, C)) & 0xFFFF)
    fids = [pyoracle.FEATURE_ID[n] for n in names]
    if kind == "holes":
        w = np.array([[-1.0, rng.gauss(-0.2, 0.2), rng.gauss(0.4, 0.3), -0.5] for _ in range(C)])
    else:
        w = np.array(_weight_sets(rng, C, names, kind))
    eng.set_features(names)
    bit = 4 if ms == "wiggle_top" else 8
    r = eng.play_games(w, G, T=T, seed=21, want_trace=True, move_set=bit)
    o = coracle.play_games(fids, w, G, T=T, seed=21, threads=16, want_trace=True, move_set=ms)
    res = r["results"]
    assert (res["pieces"] == o["pieces"]).all() and (res["lines"] == o["lines"]).all() and (res["score"] == o["score"]).all()
    assert (res["placements"] == o["placements"]).all()
    for g in range(C * G):
        n = int(res["pieces"][g])
        assert (r["trace"][g, :n] == o["trace"][g, :n]).all(), g


def test_wiggle_lookahead2_vs_coracle(eng):
    rng = random.Random(606)
    names = ["max_ht", "row_trans", "pits", "landing_height"]
    fids = [pyoracle.FEATURE_ID[n] for n in names]
    w = np.array([[-1.0, rng.gauss(-0.2, 0.2), rng.gauss(0.4, 0.3), -0.5] for _ in range(6)])
    eng.set_features(names)
    for ms, bit in (("wiggle_top", 4), ("wiggle_drop", 8)):
        r = eng.play_games(w, 2, T=16, seed=31, lookahead=2, want_trace=True, move_set=bit)
        o = coracle.play_games(fids, w, 2, T=16, seed=31, lookahead=2, threads=12, want_trace=True, move_set=ms)
        res = r["results"]
        assert (res["pieces"] == o["pieces"]).all() and (res["score"] == o["score"]).all() and (res["placements"] == o["placements"]).all()
        for g in range(12):
            n = int(res["pieces"][g])
            assert (r["trace"][g, :n] == o["trace"][g, :n]).all(), (ms, g)


def test_differential_fuzz_sample():
    """A short run of tools/fuzz.py (random feature lists, lane widths, lookahead, move sets, pressure, initial boards) —
    the long runs (6 400 configurations, 150 000 games, 2e8 placements, all bit-exact) are recorded in DESIGN.md."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz.py"), "80", "12345"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "fuzz ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
