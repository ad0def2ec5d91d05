"""Pin both CPU oracles (oracle/pyoracle.py, oracle/coracle.c) against outputs of the UNMODIFIED
reference (tests/golden/*.json, produced by tests/golden/make_golden.py in the dev container)."""
import hashlib
import random

import numpy as np
import pytest

from oracle import coracle, pyoracle

NAMES = list(pyoracle.FEATURE_NAMES)


def fval(v):
    return float.fromhex(v) if isinstance(v, str) else float(v)


def fids_weights(pairs):
    return [pyoracle.FEATURE_ID[n] for n, _ in pairs], [w for _, w in pairs]


def test_feature_order_matches_reference(golden_states):
    assert golden_states["meta"]["feature_names"] == NAMES
    assert golden_states["meta"]["piece_order"] == list(pyoracle.PIECE_NAMES)


def test_coracle_features_all_states(golden_states):
    """All 45 features, drop row, cleared count and post-clear board for every golden state (C oracle)."""
    n = 0
    for st in golden_states["states"]:
        if st["n_legal"] == 0:
            assert coracle.drop_moves(st["prev"], st["piece"]) == []
            continue
        r = coracle.eval_placement(st["prev"], st["piece"], st["rot"], st["col"])
        assert r is not None
        assert r["row"] == st["row"] and r["k"] == st["k"]
        assert list(r["post"]) == st["post"]
        want = [fval(v) for v in st["features"]]
        got = list(r["features"])
        for i, (a, b) in enumerate(zip(got, want)):
            assert a.hex() == b.hex() or (a == 0.0 and b == 0.0), (NAMES[i], a, b, st)
        if st["legal_slots"] is not None:
            mv = coracle.drop_moves(st["prev"], st["piece"])
            assert [[m[0], m[2], m[1]] for m in mv] == st["legal_slots"]
        n += 1
    assert n > 1000


def test_pyoracle_features_sample(golden_states):
    """numpy oracle on a sample of the golden states (it is ~1000x slower than the C one)."""
    rng = random.Random(5)
    states = [s for s in golden_states["states"] if s["n_legal"] > 0]
    for st in states[:4] + rng.sample(states, 150):
        root = pyoracle.Node(None, pyoracle.Grid.from_rows(st["prev"]))
        moves = pyoracle.drop_moves(root, st["piece"])
        assert len(moves) == st["n_legal"]
        assert (st["rot"], st["row"], st["col"]) in moves
        node = root.child(st["piece"], st["rot"], st["row"], st["col"])
        assert len(node.gone) == st["k"]
        assert node.grid.to_rows() == st["post"]
        got = pyoracle.all_features(node)
        for i, (a, b) in enumerate(zip(got, st["features"])):
            if isinstance(b, str):
                assert isinstance(a, float) and a.hex() == b, (NAMES[i], a, b)
            else:
                assert int(a) == b and float(a) == float(b), (NAMES[i], a, b)


def test_known_answer_board(golden_states):
    """SURVEY.md §8a-F hand-checked example: I rot1 @ col 9 clears two rows."""
    st = golden_states["states"][0]
    assert (st["piece"], st["rot"], st["row"], st["col"], st["k"]) == (0, 1, 16, 9, 2)
    f = dict(zip(NAMES, st["features"]))
    assert f["pits"] == 2 and f["pit_depth"] == 6 and f["eroded_cells"] == 2 and f["landing_height"] == 0
    assert f["col_trans"] == 13 and f["row_trans"] == 12 and f["weighted_cells"] == 82
    assert float.fromhex(f["d_mean_ht"]) == -1.5999999999999996


def test_coracle_steps(golden_steps):
    for s in golden_steps["steps"]:
        fids, w = fids_weights(s["weights"])
        for key, la in (("l1", 1), ("l2", 2)):
            want = s[key]
            if key == "l1" and want is None:
                assert coracle.best_move(s["board"], s["piece"], -1, 1, fids, w) is None
                continue
            if want is None:
                continue
            got = coracle.best_move(s["board"], s["piece"], s["next"] if la == 2 else -1, la, fids, w)
            assert (got["rot"], got["row"], got["col"], got["k"]) == (want["rot"], want["row"], want["col"], want["k"]), (key, s)
            assert got["score"].hex() == want["score"] or (got["score"] == 0.0 and float.fromhex(want["score"]) == 0.0)
            if la == 1:
                assert got["n_scored"] == want["n_legal"]
            elif want["n_leaves"]:
                assert got["n_scored"] == want["n_leaves"]


def _check_game(res, g):
    assert res["pieces"] == g["pieces"]
    assert res["lines"] == g["lines"]
    assert list(res["hist"]) == g["hist"]
    assert res["score"] == g["score"]
    trace = bytes(np.asarray(res["trace"], dtype=np.uint8).tobytes()) if not isinstance(res["trace"], bytes) else res["trace"]
    assert hashlib.sha256(trace).hexdigest()[:16] == g["trace_sha16"]
    assert trace.hex() == g["trace_hex"]


def test_coracle_games(golden_games):
    for g in golden_games["games"]:
        if g.get("tie") == "last":
            continue
        fids, w = fids_weights(g["weights"])
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        res = coracle.play_one(fids, w, seq, g["lookahead"])
        _check_game(res, g)
        assert res["placements"] == g["scorer_calls"]
        assert list(res["final"]) == g["final_board"]


def test_pyoracle_games(golden_games):
    """numpy oracle on the short golden games (whole traces)."""
    done = 0
    for g in golden_games["games"]:
        if g.get("tie") == "last" or g["scorer_calls"] > 4000:
            continue
        fids, w = fids_weights(g["weights"])
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        res = pyoracle.play_game(fids, w, seq, g["lookahead"], want_trace=True)
        _check_game(res, g)
        done += 1
    assert done >= 5


def test_hashed_piece_twins():
    rng = random.Random(3)
    counts = [0] * 7
    for _ in range(3000):
        s, g, t = rng.getrandbits(40), rng.getrandbits(24), rng.getrandbits(20)
        a = pyoracle.hashed_piece(s, g, t)
        assert a == coracle.hashed_piece(s, g, t)
        counts[a] += 1
    assert min(counts) > 300


def test_ce_golden_matches_survey(golden_ce):
    """The regenerated CE golden equals the numbers recorded in SURVEY.md §8c."""
    g0, g1 = golden_ce["ce"]["generations"]
    assert [float.fromhex(x) for x in g0["fitness"][:10]] == [0.1, 0, 0, 0, 0, 0.4, 0, 0, 0, 0]
    assert g0["mean_repr"][0] == "-8.576702798307936" and g0["std_repr"][5] == "5.635935513992518"
    assert [float.fromhex(x) for x in g1["fitness"][:10]] == [21.3, 81.1, 64.5, 57.9, 259.0, 18.2, 21.8, 51.2, 18.2, 0.6]
    assert g1["mean_repr"][0] == "-13.598518865369106" and g1["std_repr"][5] == "3.404049729891188"


def test_coracle_ce_generations(golden_ce):
    """Unmodified-reference CE run (2 generations) reproduced with pyoracle.cross_entropy driving the C oracle."""
    ce = golden_ce["ce"]
    fids = [pyoracle.FEATURE_ID[n] for n in ce["features"]]
    seqs = np.array([pyoracle.piece_sequence(ce["seed0"] + g, ce["T"]) for g in range(ce["games"])], dtype=np.uint8)
    batch = []

    def test_f(weights):
        batch.append([weights[n] for n in ce["features"]])
        return None

    log = []

    def map_f(fn, it):
        idx = list(it)
        batch.clear()
        for i in idx:
            fn(i)
        res = coracle.play_games(fids, np.array(batch), ce["games"], pieces=seqs, threads=8)
        lines = res["lines"].reshape(len(idx), ce["games"])
        fit = [int(lines[i].sum()) / ce["games"] for i in range(len(idx))]
        log.append(fit)
        return [(fit[i], i) for i in idx]

    gen = pyoracle.cross_entropy(ce["features"], ce["stdev"], ce["width"], ce["keep"], random.Random(ce["rng_seed"]), test_f, map_f=map_f)
    for want in ce["generations"]:
        mean, std = next(gen)
        assert [float(x).hex() for x in log[-1]] == want["fitness"]
        assert [float(v).hex() for v in mean.values()] == want["mean"]
        assert [float(v).hex() for v in std.values()] == want["std"]


# ------------------------------------------------------------------------------------------------
# move_with_overhang_slide(move_drop)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="session")
def golden_slide_games():
    from conftest import load_golden
    return load_golden("games_slide.json")


@pytest.fixture(scope="session")
def golden_slide_moves():
    from conftest import load_golden
    return load_golden("slide_moves.json.gz")


def test_slide_move_lists(golden_slide_moves):
    """Exact move lists (order included) of the reference's move_with_overhang_slide(move_drop)."""
    extra = 0
    for i, b in enumerate(golden_slide_moves["boards"]):
        for p, want in enumerate(b["moves"]):
            got = coracle.moves(b["board"], p, "slide")
            assert [list(m) for m in got] == want, (i, p)
            extra += len(want) - len(coracle.moves(b["board"], p, "drop"))
        if i % 10 == 0:
            node = pyoracle.Node(None, pyoracle.Grid.from_rows(b["board"]))
            for p, want in enumerate(b["moves"]):
                assert [list(m) for m in pyoracle.slide_moves(node, p)] == want
    assert extra > 2000


def test_slide_games(golden_slide_games):
    n_py = 0
    for g in golden_slide_games["games"]:
        assert g["move_gen"] == "overhang_slide"
        fids, w = fids_weights(g["weights"])
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        res = coracle.play_one(fids, w, seq, g["lookahead"], move_set="slide")
        _check_game(res, g)
        assert res["placements"] == g["scorer_calls"] and list(res["final"]) == g["final_board"]
        if g["scorer_calls"] < 1300:
            _check_game(pyoracle.play_game(fids, w, seq, g["lookahead"], want_trace=True, move_set="slide"), g)
            n_py += 1
    assert n_py >= 3


# ------------------------------------------------------------------------------------------------
# move_with_pressure(move_drop | move_with_overhang_slide(move_drop))
# ------------------------------------------------------------------------------------------------
def test_pressure_games():
    from conftest import load_golden
    n_py = 0
    for g in load_golden("games_pressure.json")["games"]:
        fids, w = fids_weights(g["weights"])
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        ms = "slide" if g["move_gen"] == "overhang_slide" else "drop"
        res = coracle.play_one(fids, w, seq, g["lookahead"], move_set=ms, pressure=g["pressure"])
        _check_game(res, g)
        assert res["placements"] == g["scorer_calls"] and list(res["final"]) == g["final_board"]
        if g["scorer_calls"] < 4000:
            _check_game(pyoracle.play_game(fids, w, seq, g["lookahead"], want_trace=True, move_set=ms, pressure=g["pressure"]), g)
            n_py += 1
    assert n_py >= 2


# ------------------------------------------------------------------------------------------------
# move_wiggle / move_with_wiggle(move_drop)
# ------------------------------------------------------------------------------------------------
def test_wiggle_move_lists():
    from conftest import load_golden
    boards = load_golden("wiggle_moves.json.gz")["boards"]
    more = 0
    for i, b in enumerate(boards):
        for p in range(7):
            assert [list(m) for m in coracle.moves(b["board"], p, "wiggle_top")] == b["top"][p], (i, p)
            assert [list(m) for m in coracle.moves(b["board"], p, "wiggle_drop")] == b["drop"][p], (i, p)
            more += len(b["top"][p]) - len(coracle.moves(b["board"], p, "drop"))
        if i % 8 == 0:
            node = pyoracle.Node(None, pyoracle.Grid.from_rows(b["board"]))
            for p in range(7):
                assert [list(m) for m in pyoracle.wiggle_moves_top(node, p)] == b["top"][p]
                assert [list(m) for m in pyoracle.wiggle_moves_drop(node, p)] == b["drop"][p]
    assert more > 100


def test_wiggle_games():
    from conftest import load_golden
    n_py = 0
    for g in load_golden("games_wiggle.json")["games"]:
        fids, w = fids_weights(g["weights"])
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        res = coracle.play_one(fids, w, seq, g["lookahead"], move_set=g["move_gen"], pressure=g["pressure"])
        _check_game(res, g)
        assert res["placements"] == g["scorer_calls"] and list(res["final"]) == g["final_board"]
        if g["scorer_calls"] < 1500:
            _check_game(pyoracle.play_game(fids, w, seq, g["lookahead"], want_trace=True, move_set=g["move_gen"], pressure=g["pressure"]), g)
            n_py += 1
    assert n_py >= 2
