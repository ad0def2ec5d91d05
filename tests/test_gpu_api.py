"""The reference-facing host API on a real GPU: simulate(), move_drop, policy_best, feature.evaluate,
cross_entropy with the batching evaluator, device-pointer (stream) calls.  Golden values come from
the unmodified reference (tests/golden)."""
import collections
import random

import numpy as np
import pytest

from oracle import coracle, pyoracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from tetris_ai_b200 import Engine
    from tetris_ai_b200 import feature
    e = Engine(0)
    feature.set_default_engine(e)
    yield e
    e.close()


def _weights(pairs):
    from tetris_ai_b200.tetris import features as F
    return collections.OrderedDict((getattr(F, n), w) for n, w in pairs)


@pytest.mark.parametrize("chunk", [4096, 37])
def test_simulate_generator_matches_golden(eng, golden_games, chunk, monkeypatch):
    from tetris_ai_b200.tetris import simulator as sim
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    monkeypatch.setattr(sim, "CHUNK", chunk)
    for g in golden_games["games"]:
        if g.get("tie") == "last" or g["pieces"] > 200:
            continue
        policy = sim.policy_best(sim.linear_scorer(_weights(g["weights"])), sim.tie_first)
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        trace = bytearray()
        last = State(None, Board(20, 10))
        n = 0
        for st in sim.simulate(last, iter(PIECES[i] for i in seq), sim.move_drop, policy, lookahead=g["lookahead"], engine=eng):
            d = st.delta
            trace += bytes([d.rot, d.row, d.col, len(d.cleared)])
            last = st
            n += 1
        assert n == g["pieces"] and trace.hex() == g["trace_hex"], (g["weights"][0], g["seed"], g["lookahead"], chunk)
        assert last.lines_cleared() == g["lines"] and last.score() == g["score"]
        assert list(last.board.rows16()) == g["final_board"]


def test_simulate_refuses_cpu_paths(eng):
    from tetris_ai_b200.tetris import simulator as sim
    from tetris_ai_b200.tetris.game import Board, State
    from tetris_ai_b200.tetris import features as F
    pol = sim.policy_best(sim.linear_scorer({F.pits: -1.0}), sim.tie_first)
    with pytest.raises(TypeError):
        next(sim.simulate(State(None, Board()), iter([]), lambda s, z: [], pol))
    with pytest.raises(TypeError):
        sim.policy_best(lambda s: 0.0, sim.tie_first)
    with pytest.raises(TypeError):
        sim.policy_best(sim.linear_scorer({F.pits: -1.0}), lambda m: m[-1])


def test_move_drop_and_evaluate_match_golden(eng, golden_states):
    from tetris_ai_b200 import feature
    from tetris_ai_b200.tetris import features as F
    from tetris_ai_b200.tetris import simulator as sim
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    names = golden_states["meta"]["feature_names"]
    feats = [getattr(F, n) for n in names]
    done = 0
    for st in golden_states["states"][:400]:
        if st["n_legal"] == 0:
            continue
        root = State(None, Board(masks=st["prev"]))
        if st.get("legal_slots") is not None:
            mv = list(sim.move_drop(root, PIECES[st["piece"]], engine=eng))
            assert [[m[0], m[2], m[1]] for m in mv] == st["legal_slots"]
        node = root.future(PIECES[st["piece"]], st["rot"], st["row"], st["col"])
        vals = feature.evaluate(node, feats, engine=eng)
        for f, want in zip(feats, st["features"]):
            got = vals[f]
            if isinstance(want, str):
                assert isinstance(got, float) and got.hex() == want, (f, got, want)
            else:
                assert got == want and isinstance(got, (int, float)), (f, got, want)
        done += 1
    assert done > 100
    # weighted form (feature.py:53): {f: value * w}
    w = feature.evaluate(node, {F.pits: 2.0, F.mean_ht: -1.0}, engine=eng)
    assert w[F.pits] == vals[F.pits] * 2.0 and w[F.mean_ht] == vals[F.mean_ht] * -1.0


def test_cross_entropy_on_gpu_matches_reference_run(eng, golden_ce):
    """BASELINE config 2 end to end: the host CE driver + batching evaluator + K2 reproduce both golden
    generations of the unmodified reference run bit for bit (fitness lists, means, standard deviations)."""
    from tetris_ai_b200.learning import PopulationEvaluator, cross_entropy
    from tetris_ai_b200.tetris import features as F
    ce = golden_ce["ce"]
    feats = [getattr(F, n) for n in ce["features"]]
    seqs = np.array([pyoracle.piece_sequence(ce["seed0"] + g, ce["T"]) for g in range(ce["games"])], dtype=np.uint8)
    ev = PopulationEvaluator(eng, feats, ce["games"], pieces=seqs)
    fits = []
    orig = ev.evaluate
    ev.evaluate = lambda w: fits.append(orig(w)) or fits[-1]
    gen = cross_entropy(feats, ce["stdev"], ce["width"], ce["keep"], random.Random(ce["rng_seed"]), ev.test_f, map_f=ev.map_f)
    for g, want in enumerate(ce["generations"]):
        mean, std = next(gen)
        assert [float(x).hex() for x in fits[g]] == want["fitness"]
        assert [float(v).hex() for v in mean.values()] == want["mean"]
        assert [float(v).hex() for v in std.values()] == want["std"]
    assert int(ev.last["results"]["placements"].sum()) == 6680307 and int(ev.last["results"]["pieces"].sum()) == 294598


def test_device_pointer_call_matches_host_call(eng):
    import torch
    from tetris_ai_b200 import _lib
    names = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
    eng.set_features(names)
    rng = np.random.default_rng(3)
    w = rng.normal(0, 5.0, size=(64, 6))
    host = eng.play_games(w, 3, T=200, seed=9)
    dev = torch.device("cuda", 0)
    d_w = torch.from_numpy(w).to(dev)
    d_res = torch.zeros(64 * 3 * 40, dtype=torch.uint8, device=dev)
    d_fit = torch.zeros(64, dtype=torch.float64, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eng.play_games_device(s.cuda_stream, 64, 3, d_w.data_ptr(), d_res.data_ptr(), 200, seed=9, fitness_ptr=d_fit.data_ptr())
    s.synchronize()
    res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=_lib.RESULT_DTYPE)
    assert (res == host["results"]).all()
    assert (d_fit.cpu().numpy() == host["fitness"]).all()
    assert eng.last_timing()["kernel_ms"] > 0


def test_error_reporting(eng):
    from tetris_ai_b200 import EngineError
    with pytest.raises(EngineError):
        eng.set_features([99])
    eng.set_features(["pits"])
    with pytest.raises(EngineError):
        eng.play_games(np.zeros((2, 1)), 1, T=10, lookahead=3)
    with pytest.raises(EngineError):
        eng.play_games(np.zeros((2, 3)), 1, T=10)


def test_simulate_with_overhang_slide(eng):
    """Host mirror: simulate(state, zoids, move_with_overhang_slide(move_drop), policy) on the device."""
    from conftest import load_golden
    from tetris_ai_b200.tetris import simulator as sim
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    for g in load_golden("games_slide.json")["games"]:
        if g["pieces"] > 120:
            continue
        policy = sim.policy_best(sim.linear_scorer(_weights(g["weights"])), sim.tie_first)
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        trace = bytearray()
        last = State(None, Board(20, 10))
        for st in sim.simulate(last, iter(PIECES[i] for i in seq), sim.move_with_overhang_slide(sim.move_drop), policy,
                               lookahead=g["lookahead"], engine=eng):
            d = st.delta
            trace += bytes([d.rot, d.row, d.col, len(d.cleared)])
            last = st
        assert trace.hex() == g["trace_hex"] and last.score() == g["score"]
    with pytest.raises(TypeError):
        sim.move_with_overhang_slide(lambda s, z: [])


@pytest.mark.parametrize("chunk", [4096, 23])
def test_simulate_with_pressure_chunked(eng, chunk, monkeypatch):
    """move_with_pressure depends on the level, i.e. on the lines cleared so far: chunked simulate() must carry them."""
    from conftest import load_golden
    from tetris_ai_b200.tetris import simulator as sim
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    monkeypatch.setattr(sim, "CHUNK", chunk)
    for g in load_golden("games_pressure.json")["games"]:
        if g["scorer_calls"] > 6100:
            continue
        inner = sim.move_with_overhang_slide(sim.move_drop) if g["move_gen"] == "overhang_slide" else sim.move_drop
        gen = sim.move_with_pressure(inner, **g["pressure"])
        policy = sim.policy_best(sim.linear_scorer(_weights(g["weights"])), sim.tie_first)
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        trace = bytearray()
        last = State(None, Board(20, 10))
        for st in sim.simulate(last, iter(PIECES[i] for i in seq), gen, policy, lookahead=g["lookahead"], engine=eng):
            d = st.delta
            trace += bytes([d.rot, d.row, d.col, len(d.cleared)])
            last = st
        assert trace.hex() == g["trace_hex"] and last.score() == g["score"], (g["seed"], chunk)


def test_evaluate_on_slide_states(eng):
    """feature.evaluate on states produced by overhang slides (explicit-row K3 path) equals the C oracle's features."""
    from conftest import load_golden
    from tetris_ai_b200 import feature
    from tetris_ai_b200.tetris import features as F
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    feats = [getattr(F, n) for n in pyoracle.FEATURE_NAMES]
    checked = 0
    for b in load_golden("slide_moves.json.gz")["boards"][:40]:
        for p, moves in enumerate(b["moves"]):
            base = set(coracle.moves(b["board"], p, "drop"))
            for rot, row, col in moves:
                if (rot, row, col) in base:
                    continue
                root = State(None, Board(masks=b["board"]))
                node = root.future(PIECES[p], rot, row, col)
                vals = feature.evaluate(node, feats, engine=eng)
                onode = pyoracle.Node(None, pyoracle.Grid.from_rows(b["board"])).child(p, rot, row, col)
                want = pyoracle.all_features(onode)
                for f, wv in zip(feats, want):
                    assert float(vals[f]) == float(wv), (f, vals[f], wv)
                checked += 1
                if checked >= 60:
                    return
    assert checked > 10


def test_cross_entropy_six_generations_vs_c_oracle(eng):
    """A longer CE run (later generations: most games reach the cap): GPU evaluation vs the plain-C oracle evaluation must give
    bit-identical means / standard deviations generation after generation (one mismatched placement anywhere would diverge)."""
    from tetris_ai_b200.learning import PopulationEvaluator, cross_entropy
    from tetris_ai_b200.tetris import features as F
    names = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
    feats = [getattr(F, n) for n in names]
    fids = [pyoracle.FEATURE_ID[n] for n in names]
    G, T, width, keep = 6, 500, 60, 8
    seqs = np.array([pyoracle.piece_sequence(2000 + g, T) for g in range(G)], dtype=np.uint8)
    ev = PopulationEvaluator(eng, feats, G, pieces=seqs)
    gpu = cross_entropy(feats, 10.0, width, keep, random.Random(7), ev.test_f, map_f=ev.map_f)

    batch = []

    def test_f(wd):
        batch.append(list(wd.values()))

    def map_f(fn, it):
        idx = list(it)
        batch.clear()
        for i in idx:
            fn(i)
        res = coracle.play_games(fids, np.array(batch), G, pieces=seqs, threads=16)
        fit = res["lines"].reshape(len(idx), G).sum(axis=1) / G
        return [(float(fit[i]), i) for i in idx]

    cpu = cross_entropy(feats, 10.0, width, keep, random.Random(7), test_f, map_f=map_f)
    for g in range(6):
        (m1, s1), (m2, s2) = next(gpu), next(cpu)
        assert [float(v).hex() for v in m1.values()] == [float(v).hex() for v in m2.values()], g
        assert [float(v).hex() for v in s1.values()] == [float(v).hex() for v in s2.values()], g
    assert float(np.max(ev.last["fitness"])) > 50.0           # the population has learned to clear lines


def test_simulate_with_wiggle(eng):
    from conftest import load_golden
    from tetris_ai_b200.tetris import simulator as sim
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    for g in load_golden("games_wiggle.json")["games"][-7:]:
        gen = sim.move_wiggle if g["move_gen"] == "wiggle_top" else sim.move_with_wiggle(sim.move_drop)
        if g["pressure"] is not None:
            gen = sim.move_with_pressure(gen, **g["pressure"])
        policy = sim.policy_best(sim.linear_scorer(_weights(g["weights"])), sim.tie_first)
        seq = pyoracle.piece_sequence(g["seed"], g["T"])
        trace = bytearray()
        last = State(None, Board(20, 10))
        for st in sim.simulate(last, iter(PIECES[i] for i in seq), gen, policy, lookahead=g["lookahead"], engine=eng):
            d = st.delta
            trace += bytes([d.rot, d.row, d.col, len(d.cleared)])
            last = st
        assert trace.hex() == g["trace_hex"] and last.score() == g["score"]


def test_graph_replay_then_direct_calls(eng):
    """A CUDA-graph replay (repeated evaluator call on pinned buffers) followed by ordinary calls: no stale CUDA error,
    timings stay sane, results identical to the direct path."""
    from tetris_ai_b200.learning import PopulationEvaluator
    names = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
    rng = np.random.default_rng(8)
    ev = PopulationEvaluator(eng, names, 4, T=120, seed=3)
    w = rng.normal(0, 5, size=(50, 6))
    fits = [ev.evaluate(w).copy() for _ in range(3)]            # capture, replay, replay
    assert (fits[0] == fits[1]).all() and (fits[1] == fits[2]).all()
    t = eng.last_timing()
    assert t["total_ms"] > 0.0
    w2 = rng.normal(0, 5, size=(50, 6))
    f2 = ev.evaluate(w2).copy()                                  # replay with new weights in the same pinned buffer
    eng.set_features(names)
    direct = eng.play_games(w2, 4, T=120, seed=3)                # pageable buffers: direct path right after a replay
    assert (direct["fitness"] == f2).all()
    assert eng.last_timing()["kernel_ms"] > 0.0
