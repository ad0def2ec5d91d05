"""The CUDA core (tetris_core.cuh) compiled for the HOST, checked against the golden fixtures and
against the plain-C oracle on thousands of seeded games.  Runs without a GPU; the same checks are
repeated through the real kernels in test_gpu_parity.py (-m gpu)."""
import random

import numpy as np
import pytest

from oracle import coracle, pyoracle
import emu_lib

NAMES = list(pyoracle.FEATURE_NAMES)
W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]


def fval(v):
    return float.fromhex(v) if isinstance(v, str) else float(v)


def same_double(a, b):
    return a.hex() == b.hex() or (a == 0.0 and b == 0.0)


def test_emu_features_all_golden_states(golden_states):
    n = 0
    for st in golden_states["states"]:
        if st["n_legal"] == 0:
            continue
        r = emu_lib.eval_placement(st["prev"], st["piece"], st["rot"], st["col"])
        assert r["legal"] == 1 and r["row"] == st["row"] and r["k"] == st["k"]
        assert list(r["post"]) == st["post"]
        for i, (a, b) in enumerate(zip(r["features"], st["features"])):
            assert same_double(float(a), fval(b)), (NAMES[i], float(a), b, st)
        n += 1
    assert n > 7000


def test_emu_legality_all_slots(golden_states):
    """Every (rot, col) slot: legal exactly when move_drop yields it, with the same drop row."""
    nrot = [2, 4, 4, 4, 1, 2, 2]
    checked = 0
    for st in golden_states["states"]:
        if st.get("legal_slots") is None:
            continue
        want = {(r, c): row for r, c, row in st["legal_slots"]}
        for rot in range(nrot[st["piece"]]):
            for col in range(10):
                r = emu_lib.eval_placement(st["prev"], st["piece"], rot, col)
                if (rot, col) in want:
                    assert r["legal"] == 1 and r["row"] == want[(rot, col)]
                else:
                    assert r["legal"] in (0, -1)
        checked += 1
    assert checked > 200


def test_emu_golden_games(golden_games):
    for g in golden_games["games"]:
        if g.get("tie") == "last":
            continue
        fids = [pyoracle.FEATURE_ID[n] for n, _ in g["weights"]]
        w = [x for _, x in g["weights"]]
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        for mode in (-1, 0):
            res = emu_lib.play_one(fids, w, seq, lookahead=g["lookahead"], mode=mode)
            assert res["pieces"] == g["pieces"] and res["lines"] == g["lines"] and res["hist"] == g["hist"]
            assert res["score"] == g["score"] and res["placements"] == g["scorer_calls"]
            assert res["trace"].tobytes().hex() == g["trace_hex"]
            assert list(res["final"]) == g["final_board"]


def test_emu_golden_steps(golden_steps):
    for s in golden_steps["steps"]:
        fids = [pyoracle.FEATURE_ID[n] for n, _ in s["weights"]]
        w = [x for _, x in s["weights"]]
        for key, la in (("l1", 1), ("l2", 2)):
            want = s[key]
            if want is None and key == "l2":
                continue
            res = emu_lib.play_one(fids, w, [s["piece"], s["next"]], lookahead=la, max_moves=1, init_rows=s["board"])
            if want is None:
                assert res["pieces"] == 0
                continue
            assert res["pieces"] == 1
            assert tuple(int(x) for x in res["trace"][0]) == (want["rot"], want["row"], want["col"], want["k"])
            assert same_double(float(res["scores"][0]), float.fromhex(want["score"]))


def _random_weight_sets(rng, n):
    out = []
    for i in range(n):
        kind = i % 4
        if kind == 0:
            out.append((W6_NAMES, [rng.gauss(v, 1.0) for v in W6_VALS]))
        elif kind == 1:
            out.append((W6_NAMES, [rng.gauss(0, 10.0) for _ in W6_VALS]))
        elif kind == 2:
            names = rng.sample(NAMES, rng.randint(1, 12))
            out.append((names, [rng.gauss(0, 3.0) for _ in names]))
        else:
            out.append((NAMES, [W6_VALS[W6_NAMES.index(n)] + rng.gauss(0, 0.3) if n in W6_NAMES else rng.gauss(0, 0.3) for n in NAMES]))
    return out


@pytest.mark.parametrize("lookahead,ngames,T", [(1, 160, 400), (2, 24, 40)])
def test_emu_vs_coracle_random_games(lookahead, ngames, T):
    """Seeded random weight vectors (w6-like, wide, random subsets, all 45) x hashed piece streams."""
    rng = random.Random(100 + lookahead)
    for gi, (names, w) in enumerate(_random_weight_sets(rng, ngames)):
        fids = [pyoracle.FEATURE_ID[n] for n in names]
        seq = [pyoracle.hashed_piece(77, gi, t) for t in range(T)]
        a = coracle.play_one(fids, w, seq, lookahead)
        b = emu_lib.play_one(fids, w, seq, lookahead=lookahead)
        assert (a["pieces"], a["lines"], a["hist"], a["score"], a["placements"]) == \
               (b["pieces"], b["lines"], b["hist"], b["score"], b["placements"]), (gi, names)
        assert a["trace"].tobytes() == b["trace"].tobytes(), (gi, names)
        assert list(a["final"]) == list(b["final"])


def test_emu_hashed_piece_twin():
    rng = random.Random(9)
    for _ in range(2000):
        s, g, t = rng.getrandbits(48), rng.getrandbits(30), rng.getrandbits(20)
        assert emu_lib.lib().emu_hashed_piece(s, g, t) == pyoracle.hashed_piece(s, g, t)


def test_emu_fast_w6_path_vs_coracle(golden_games):
    """The w32 kernel's arithmetic-heights / incremental-cells fast path for the w6 set (emu mode 3)."""
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    for g in golden_games["games"]:
        if [n for n, _ in g["weights"]] != W6_NAMES or g["lookahead"] != 1 or g.get("tie") == "last":
            continue
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        res = emu_lib.play_one(fids, [x for _, x in g["weights"]], seq, mode=3)
        assert res["trace"].tobytes().hex() == g["trace_hex"] and res["score"] == g["score"]
    rng = random.Random(321)
    for gi in range(300):
        w = [rng.gauss(v, 1.5) for v in W6_VALS] if gi % 2 else [rng.gauss(0, 10.0) for _ in W6_VALS]
        seq = [pyoracle.hashed_piece(5, gi, t) for t in range(500)]
        a = coracle.play_one(fids, w, seq, 1)
        b = emu_lib.play_one(fids, w, seq, mode=3)
        assert (a["pieces"], a["lines"], a["hist"], a["score"], a["placements"]) == (b["pieces"], b["lines"], b["hist"], b["score"], b["placements"]), gi
        assert a["trace"].tobytes() == b["trace"].tobytes() and list(a["final"]) == list(b["final"])
    # from a non-empty initial board
    init = [0] * 16 + [0x0F0, 0x3CF, 0x1FF, 0x3FE]
    for gi in range(20):
        w = [rng.gauss(v, 1.0) for v in W6_VALS]
        seq = [pyoracle.hashed_piece(6, gi, t) for t in range(200)]
        a = coracle.play_one(fids, w, seq, 1, init_rows=init)
        b = emu_lib.play_one(fids, w, seq, mode=3, init_rows=init)
        assert a["trace"].tobytes() == b["trace"].tobytes() and a["score"] == b["score"]


# ------------------------------------------------------------------------------------------------
# move_with_overhang_slide(move_drop) in the CUDA core (host build)
# ------------------------------------------------------------------------------------------------
def test_emu_slide_move_lists():
    from conftest import load_golden
    boards = load_golden("slide_moves.json.gz")["boards"]
    for i, b in enumerate(boards):
        for p, want in enumerate(b["moves"]):
            assert [list(m) for m in emu_lib.moves(b["board"], p, 1)] == want, (i, p)


def test_emu_slide_games():
    from conftest import load_golden
    for g in load_golden("games_slide.json")["games"]:
        fids = [pyoracle.FEATURE_ID[n] for n, _ in g["weights"]]
        w = [x for _, x in g["weights"]]
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        for mode in (-1, 0):
            res = emu_lib.play_one(fids, w, seq, lookahead=g["lookahead"], mode=mode, move_set=1)
            assert res["pieces"] == g["pieces"] and res["lines"] == g["lines"] and res["score"] == g["score"]
            assert res["placements"] == g["scorer_calls"] and res["trace"].tobytes().hex() == g["trace_hex"]


@pytest.mark.parametrize("lookahead,ngames,T", [(1, 120, 300), (2, 10, 30)])
def test_emu_slide_vs_coracle_random(lookahead, ngames, T):
    rng = random.Random(500 + lookahead)
    slid = 0
    for gi in range(ngames):
        if gi % 3 == 0:
            names, w = W6_NAMES, [rng.gauss(v, 2.0) for v in W6_VALS]
        elif gi % 3 == 1:
            names, w = ["max_ht", "row_trans", "pits", "landing_height"], [-1.0, rng.gauss(-0.2, 0.2), rng.gauss(0.4, 0.3), -0.5]
        else:
            names = rng.sample(NAMES, rng.randint(2, 10)); w = [rng.gauss(0, 2.0) for _ in names]
        fids = [pyoracle.FEATURE_ID[n] for n in names]
        seq = [pyoracle.hashed_piece(88, gi, t) for t in range(T)]
        a = coracle.play_one(fids, w, seq, lookahead, move_set="slide")
        b = emu_lib.play_one(fids, w, seq, lookahead=lookahead, move_set=1)
        assert (a["pieces"], a["lines"], a["hist"], a["score"], a["placements"]) == (b["pieces"], b["lines"], b["hist"], b["score"], b["placements"]), gi
        assert a["trace"].tobytes() == b["trace"].tobytes(), gi
        d = coracle.play_one(fids, w, seq, lookahead, move_set="drop")
        slid += a["trace"].tobytes() != d["trace"].tobytes()
    assert slid > ngames // 10          # the slide move set really changes games


# ------------------------------------------------------------------------------------------------
# move_with_pressure in the CUDA core (host build)
# ------------------------------------------------------------------------------------------------
def test_emu_pressure_games():
    from conftest import load_golden
    for g in load_golden("games_pressure.json")["games"]:
        fids = [pyoracle.FEATURE_ID[n] for n, _ in g["weights"]]
        w = [x for _, x in g["weights"]]
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        ms = 1 if g["move_gen"] == "overhang_slide" else 0
        res = emu_lib.play_one(fids, w, seq, lookahead=g["lookahead"], move_set=ms, pressure=g["pressure"])
        assert res["pieces"] == g["pieces"] and res["lines"] == g["lines"] and res["score"] == g["score"]
        assert res["placements"] == g["scorer_calls"] and res["trace"].tobytes().hex() == g["trace_hex"]


def test_emu_pressure_vs_coracle_random():
    rng = random.Random(901)
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    for gi in range(60):
        w = [rng.gauss(v, 1.0) for v in W6_VALS]
        pr = {"avg_lat": rng.uniform(120, 320), "resp_time": rng.uniform(40, 120), "move_eff": rng.uniform(1.0, 1.5), "two_but_rot": bool(gi & 1)}
        seq = [pyoracle.hashed_piece(91, gi, t) for t in range(600)]
        ms = gi % 2
        la = 2 if gi % 10 == 0 else 1
        if la == 2:
            seq = seq[:40]
        a = coracle.play_one(fids, w, seq, la, move_set=ms, pressure=pr)
        b = emu_lib.play_one(fids, w, seq, lookahead=la, move_set=ms, pressure=pr)
        assert (a["pieces"], a["lines"], a["score"], a["placements"]) == (b["pieces"], b["lines"], b["score"], b["placements"]), gi
        assert a["trace"].tobytes() == b["trace"].tobytes(), gi


# ------------------------------------------------------------------------------------------------
# move_wiggle / move_with_wiggle(move_drop) in the CUDA core (host build)
# ------------------------------------------------------------------------------------------------
def test_emu_wiggle_move_lists():
    from conftest import load_golden
    for i, b in enumerate(load_golden("wiggle_moves.json.gz")["boards"]):
        for p in range(7):
            assert [list(m) for m in emu_lib.moves(b["board"], p, 4)] == b["top"][p], (i, p)
            assert [list(m) for m in emu_lib.moves(b["board"], p, 8)] == b["drop"][p], (i, p)


def test_emu_wiggle_games():
    from conftest import load_golden
    for g in load_golden("games_wiggle.json")["games"]:
        fids = [pyoracle.FEATURE_ID[n] for n, _ in g["weights"]]
        w = [x for _, x in g["weights"]]
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        ms = 4 if g["move_gen"] == "wiggle_top" else 8
        res = emu_lib.play_one(fids, w, seq, lookahead=g["lookahead"], move_set=ms, pressure=g["pressure"])
        assert res["pieces"] == g["pieces"] and res["lines"] == g["lines"] and res["score"] == g["score"]
        assert res["placements"] == g["scorer_calls"] and res["trace"].tobytes().hex() == g["trace_hex"]


def test_emu_wiggle_vs_coracle_random():
    rng = random.Random(77)
    for gi in range(40):
        if gi % 2:
            names, w = ["max_ht", "row_trans", "pits", "landing_height"], [-1.0, rng.gauss(-0.2, 0.2), rng.gauss(0.4, 0.3), -0.5]
        else:
            names, w = W6_NAMES, [rng.gauss(v, 2.0) for v in W6_VALS]
        fids = [pyoracle.FEATURE_ID[n] for n in names]
        seq = [pyoracle.hashed_piece(55, gi, t) for t in range(200)]
        ms = ("wiggle_top", 4) if gi % 4 < 2 else ("wiggle_drop", 8)
        a = coracle.play_one(fids, w, seq, 1, move_set=ms[0])
        b = emu_lib.play_one(fids, w, seq, move_set=ms[1])
        assert (a["pieces"], a["lines"], a["score"], a["placements"]) == (b["pieces"], b["lines"], b["score"], b["placements"]), gi
        assert a["trace"].tobytes() == b["trace"].tobytes(), gi


def test_emu_wiggle_set_equals_list_as_set():
    """The bit-parallel flood fill (wiggle_reach) yields exactly the moves of the reference's recursion, as a set."""
    from conftest import load_golden
    import random as _r
    for i, b in enumerate(load_golden("wiggle_moves.json.gz")["boards"]):
        for p in range(7):
            assert emu_lib.wiggle_set(b["board"], p, True) == sorted(tuple(m) for m in b["top"][p]), (i, p)
            assert emu_lib.wiggle_set(b["board"], p, False) == sorted(tuple(m) for m in b["drop"][p]), (i, p)
    rng = _r.Random(4)
    for _ in range(300):                       # random (unreachable, cave-rich) boards against the C oracle's recursion
        hmax = rng.randint(0, 19)
        rows = [0] * 20
        for r in range(20 - hmax, 20):
            v = rng.getrandbits(10) & rng.getrandbits(10) | (rng.getrandbits(10) if rng.random() < 0.5 else 0)
            rows[r] = v if v != 0x3FF else 0x3FE
        p = rng.randrange(7)
        assert emu_lib.wiggle_set(rows, p, True) == sorted(coracle.moves(rows, p, "wiggle_top"))
        assert emu_lib.wiggle_set(rows, p, False) == sorted(coracle.moves(rows, p, "wiggle_drop"))
        # the tie break: the early-out search returns the first listed move of any subset
        assert emu_lib.wiggle_stop_check(rows, p, True, rng.getrandbits(32)) == 0
        assert emu_lib.wiggle_stop_check(rows, p, False, rng.getrandbits(32)) == 0


def test_emu_window_evaluation_equals_general_w6(golden_states):
    """win_eval (window-local w6 evaluation used by the lockstep kernel) == w6_pre + w6_post, key for key, for every slot of
    every piece on the golden boards and on random boards (flat, jagged, holey, nearly full rows), several weight sets."""
    import ctypes
    L = emu_lib.lib()
    L.emu_win_check.restype = ctypes.c_int
    rng = np.random.default_rng(11)
    boards = [np.array(st["prev"], dtype=np.uint16) for st in golden_states["states"][::9]]
    for i in range(1500):
        rows = np.zeros(20, dtype=np.uint16)
        top = int(rng.integers(0, 19))
        dens = float(rng.uniform(0.35, 0.97))
        for r in range(20 - top, 20):
            m = 0
            for k in range(10):
                if rng.random() < dens:
                    m |= 1 << k
            rows[r] = m if m != 0x3FF else m & ~(1 << int(rng.integers(0, 10)))      # boards in play have no full rows
        boards.append(rows)
    wsets = [np.array([-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]),
             np.array([1.0, -1.0, 1.0, 1.0, 1.0, 1.0]), rng.normal(0, 10, 6), np.array([0.0, 0.0, -1.0, 0.0, 0.0, 0.0]),
             np.array([0.0, 0.0, 0.0, -1.0, 0.0, 0.0]), np.array([0.0, 0.0, 0.0, 0.0, -1.0, 0.0]), np.array([0.0, 0.0, 0.0, 0.0, 0.0, -1.0])]
    total = clearing = 0
    for bi, rows in enumerate(boards):
        w = np.ascontiguousarray(wsets[bi % len(wsets)], dtype=np.float64)
        ncmp, nclr = ctypes.c_int(0), ctypes.c_int(0)
        bad = L.emu_win_check(rows.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ncmp), ctypes.byref(nclr))
        assert bad == 0, (bi, [int(x) for x in rows], bad)
        total += ncmp.value
        clearing += nclr.value
    assert total > 100000 and clearing > 1000
