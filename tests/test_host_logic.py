"""CPU-only tests of the host side: C-ABI surface, host mirrors of the reference's interfaces,
the cross-entropy driver, the evaluator adapters and the sharded (world_size 2, gloo) path."""
import collections
import os
import pickle
import random
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from tetris_ai_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tetris_b200.h")).read()
    declared = set(re.findall(r"\b(tb200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.tb200_version()


def test_c_abi_struct_layout_matches_header():
    import ctypes
    from tetris_ai_b200 import _lib
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "tetris_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(tb200_play_args), sizeof(tb200_game_result), sizeof(tb200_timing), offsetof(tb200_play_args, results), offsetof(tb200_play_args, fitness), offsetof(tb200_game_result, score), sizeof(tb200_move_opts), offsetof(tb200_move_opts, lines));return 0;}\n'
    exe = os.path.join(ROOT, "tests", "emu", "abi_probe")
    subprocess.run(["gcc", "-x", "c", "-I", os.path.join(ROOT, "include"), "-o", exe, "-"], input=src.encode(), check=True)
    sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes[0] == ctypes.sizeof(_lib.PlayArgs)
    assert sizes[1] == _lib.RESULT_DTYPE.itemsize == 40
    assert sizes[2] == ctypes.sizeof(_lib.Timing)
    assert sizes[3] == _lib.PlayArgs.results.offset and sizes[4] == _lib.PlayArgs.fitness.offset
    assert sizes[5] == _lib.RESULT_DTYPE.fields["score"][1]
    assert sizes[6] == ctypes.sizeof(_lib.MoveOpts) and sizes[7] == _lib.MoveOpts.lines.offset


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tetris_ai_b200 import Engine, EngineError
    with pytest.raises(EngineError):
        Engine(0)


def test_feature_handles_match_reference_order():
    from tetris_ai_b200.tetris import features as F
    from oracle import pyoracle
    assert list(F.FEATURE_NAMES) == list(pyoracle.FEATURE_NAMES)
    assert F.landing_height.id == 37 and F.cuml_eroded.id == 44 and F.mean_ht.__name__ == "mean_ht"
    assert not F.is_helper(F.pits)
    d = {F.pits: 1.0, F.max_ht: 2.0}
    assert list(d)[0] is F.pits


def test_game_model_replay_matches_oracle(golden_games):
    """State.future / score / level / pickling replay the golden traces to the golden results."""
    from tetris_ai_b200.tetris.game import Board, State, PIECES, zoids
    from oracle import pyoracle
    assert [z.name for z in zoids.classic] == list("ITLJOZS")
    for p, z in enumerate(PIECES):
        for rot in range(len(z)):
            assert (z[rot] == pyoracle.PIECES[p][rot]).all()
    g = golden_games["games"][0]
    seq = pyoracle.piece_sequence(g["seed"], g["T"])
    tr = bytes.fromhex(g["trace_hex"])
    st = State(None, Board(20, 10))
    for t in range(g["pieces"]):
        rot, row, col, k = tr[4 * t: 4 * t + 4]
        st = st.future(PIECES[seq[t]], rot, row, col)
        assert len(st.delta.cleared) == k
    assert st.lines_cleared() == g["lines"] and st.score() == g["score"] and st.level() == g["lines"] // 10
    assert list(st.board.rows16()) == g["final_board"]
    assert [st.lines_cleared(k) for k in (1, 2, 3, 4)] == g["hist"]
    clone = pickle.loads(pickle.dumps(st))
    assert clone.board == st.board and clone.score() == st.score() and clone.delta.row == st.delta.row


def _synthetic_test_f(weights):
    vals = list(weights.values())
    return -sum((v - 3.0 * (i + 1)) ** 2 for i, v in enumerate(vals))


def test_cross_entropy_matches_reference_driver():
    """Our host driver vs the UNMODIFIED reference driver (only where /root/reference exists)."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this box")
    import collections.abc
    collections.Mapping = collections.abc.Mapping
    sys.path.insert(0, ref)
    try:
        from cogworks import learning as ref_learning
    finally:
        sys.path.remove(ref)
    from tetris_ai_b200.learning import cross_entropy
    for feats, stdev in ((["a", "b", "c"], 5.0), (collections.OrderedDict(a=1.0, b=-2.0), {"a": 3.0, "b": 0.5})):
        g1 = ref_learning.cross_entropy(feats, stdev, 30, 5, random.Random(9), _synthetic_test_f, noise_f=lambda s: s + 0.01)
        g2 = cross_entropy(feats, stdev, 30, 5, random.Random(9), _synthetic_test_f, noise_f=lambda s: s + 0.01)
        for _ in range(4):
            (m1, s1), (m2, s2) = next(g1), next(g2)
            assert list(m1.items()) == list(m2.items()) and list(s1.items()) == list(s2.items())


def test_cross_entropy_matches_oracle_driver():
    from oracle import pyoracle
    from tetris_ai_b200.learning import cross_entropy
    g1 = pyoracle.cross_entropy(["a", "b"], 4.0, 20, 4, random.Random(1), _synthetic_test_f)
    g2 = cross_entropy(["a", "b"], 4.0, 20, 4, random.Random(1), _synthetic_test_f)
    for _ in range(3):
        assert next(g1) == next(g2)


class _FakeEngine:
    """Stands in for Engine in CPU tests: same play_games signature, evaluated by the C oracle."""

    def set_features(self, names):
        from oracle import pyoracle
        self.fids = [pyoracle.FEATURE_ID[n] for n in names]

    def play_games(self, weights, G, pieces=None, T=None, seed=0, lookahead=1, lanes=0, **kw):
        from oracle import coracle
        o = coracle.play_games(self.fids, weights, G, pieces=pieces, T=T, seed=seed, lookahead=lookahead, threads=8)
        C = len(weights)
        return {"fitness": o["lines"].reshape(C, G).sum(axis=1) / G, "results": o}


def test_population_evaluator_drives_ce_to_golden(golden_ce):
    """The batching test_f/map_f adapter reproduces generation 0 of the golden CE run (evaluation by the C oracle here;
    the same adapter with the real Engine is checked in test_gpu_ce.py)."""
    from tetris_ai_b200.learning import PopulationEvaluator, cross_entropy
    from tetris_ai_b200.tetris import features as F
    from oracle import pyoracle
    ce = golden_ce["ce"]
    feats = [getattr(F, n) for n in ce["features"]]
    seqs = np.array([pyoracle.piece_sequence(ce["seed0"] + g, ce["T"]) for g in range(ce["games"])], dtype=np.uint8)
    ev = PopulationEvaluator(_FakeEngine(), feats, ce["games"], pieces=seqs)
    gen = cross_entropy(feats, ce["stdev"], ce["width"], ce["keep"], random.Random(ce["rng_seed"]), ev.test_f, map_f=ev.map_f)
    mean, std = next(gen)
    assert [float(v).hex() for v in mean.values()] == ce["generations"][0]["mean"]
    assert [float(v).hex() for v in std.values()] == ce["generations"][0]["std"]
    assert [f.__name__ for f in mean] == ce["features"]


def test_shard_bounds_cover_everything():
    from tetris_ai_b200.learning import shard_bounds
    for n in (1, 7, 100, 10000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_sharded_evaluator_world2_gloo(tmp_path):
    """world_size 2 over gloo: both ranks end with the full, identical fitness vector and CE refit."""
    import json
    script = tmp_path / "w2.py"
    script.write_text('''
import os, sys, random, json
sys.path.insert(0, %r)
import numpy as np
import torch.distributed as dist
from tetris_ai_b200.learning import ShardedEvaluator, cross_entropy
dist.init_process_group("gloo")
calls = []
def local_eval(w):
    calls.append(len(w))
    return -((w - np.arange(1, w.shape[1] + 1) * 2.0) ** 2).sum(axis=1)
ev = ShardedEvaluator(local_eval)
gen = cross_entropy(["a", "b", "c"], 5.0, 21, 4, random.Random(3), ev.test_f, map_f=ev.map_f)
out = [next(gen) for _ in range(3)]
with open(os.path.join(%r, "rank%%d.json" %% dist.get_rank()), "w") as f:
    json.dump({"refits": [[[float(x).hex() for x in m.values()], [float(x).hex() for x in s.values()]] for m, s in out], "calls": calls}, f)
dist.destroy_process_group()
''' % (ROOT, str(tmp_path)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29571", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    r0 = json.load(open(tmp_path / "rank0.json"))
    r1 = json.load(open(tmp_path / "rank1.json"))
    assert r0["refits"] == r1["refits"]                       # bit-identical refits on both ranks
    assert r0["calls"] == [11, 11, 11] and r1["calls"] == [10, 10, 10]      # 21 candidates split 11 / 10
    # and bit-identical to the single-process run
    from tetris_ai_b200.learning import cross_entropy

    def test_f(w):
        v = np.array(list(w.values()))[None, :]
        return float(-((v - np.arange(1, v.shape[1] + 1) * 2.0) ** 2).sum(axis=1)[0])
    gen = cross_entropy(["a", "b", "c"], 5.0, 21, 4, random.Random(3), test_f)
    want = [[[float(x).hex() for x in m.values()], [float(x).hex() for x in s.values()]] for m, s in (next(gen) for _ in range(3))]
    assert r0["refits"] == want


def test_sharded_evaluator_detects_desync_and_bad_keys(tmp_path):
    """world_size 2 over gloo: ranks that sample different populations (different seeds) must raise, not refit silently;
    a weights mapping whose key order differs from the evaluator's feature order must raise on every rank."""
    script = tmp_path / "w2bad.py"
    script.write_text('''
import os, sys, random
sys.path.insert(0, %r)
import numpy as np
import torch.distributed as dist
from tetris_ai_b200.learning import ShardedEvaluator, cross_entropy
dist.init_process_group("gloo")
ev = ShardedEvaluator(lambda w: w.sum(axis=1), names=["a", "b", "c"])
try:
    ev.test_f({"a": 1.0, "c": 2.0, "b": 3.0})
    print("KEYS_NOT_CHECKED")
except ValueError:
    print("KEYS_OK")
gen = cross_entropy(["a", "b", "c"], 5.0, 8, 2, random.Random(3 + dist.get_rank()), ev.test_f, map_f=ev.map_f)
try:
    next(gen)
    print("DESYNC_NOT_DETECTED")
except RuntimeError as e:
    print("DESYNC_OK" if "different weight batches" in str(e) else "OTHER %%s" %% e)
dist.destroy_process_group()
''' % (ROOT,))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29573", str(script)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("KEYS_OK") == 2 and out.stdout.count("DESYNC_OK") == 2, out.stdout


def test_unmodified_reference_driver_with_evaluator_adapter(golden_ce):
    """the UNMODIFIED cogworks.learning.cross_entropy drives the batching adapter with the
    reference's own feature objects as mapping keys (evaluation by the C oracle here; by the GPU in test_gpu_api)."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present on this box")
    import collections.abc
    collections.Mapping = collections.abc.Mapping
    sys.path.insert(0, ref)
    try:
        from cogworks import learning as ref_learning
        from cogworks.tetris import features as ref_features
    finally:
        sys.path.remove(ref)
    from tetris_ai_b200.learning import PopulationEvaluator
    from oracle import pyoracle
    ce = golden_ce["ce"]
    feats = [getattr(ref_features, n) for n in ce["features"]]
    seqs = np.array([pyoracle.piece_sequence(ce["seed0"] + g, ce["T"]) for g in range(ce["games"])], dtype=np.uint8)
    ev = PopulationEvaluator(_FakeEngine(), feats, ce["games"], pieces=seqs)
    gen = ref_learning.cross_entropy(feats, ce["stdev"], ce["width"], ce["keep"], random.Random(ce["rng_seed"]), ev.test_f, map_f=ev.map_f)
    mean, std = next(gen)
    assert [float(v).hex() for v in mean.values()] == ce["generations"][0]["mean"]
    assert [float(v).hex() for v in std.values()] == ce["generations"][0]["std"]


class _SimFakeEngine:
    """Engine stand-in for simulate(): one game per call, played by the C oracle (a checker, fine in tests)."""

    def __init__(self):
        self.launches = 0

    def set_features(self, feats):
        from oracle import pyoracle
        self.fids = [pyoracle.FEATURE_ID[getattr(f, "__name__", f)] for f in feats]

    def play_games(self, weights, G, pieces=None, lookahead=1, max_moves=0, init_rows=None, init_lines=0, **kw):
        from oracle import coracle
        self.launches += 1
        o = coracle.play_one(self.fids, weights[0], pieces[0], lookahead=lookahead, init_rows=init_rows[0], init_lines=init_lines)
        n = min(o["pieces"], max_moves) if max_moves else o["pieces"]
        trace = np.zeros((1, pieces.shape[1], 4), dtype=np.uint8)
        trace[0, :n] = o["trace"][:n]
        return {"results": {"pieces": np.array([n])}, "trace": trace}


@pytest.mark.parametrize("mode", ["stream", "chunk1"])
def test_simulate_piece_consumption_matches_reference(mode):
    """the reference pulls exactly `lookahead` pieces per move from zoid_gen (simulator.py:156).  With a PieceStream
    (push-back) or chunk=1 a generator shared between successive games is left exactly where the reference leaves it
    (counts and the next game's first piece from tests/golden/r2_extra.json, generated from the unmodified reference)."""
    import collections
    from conftest import load_golden
    from oracle import pyoracle
    from tetris_ai_b200.tetris import features as F
    from tetris_ai_b200.tetris import simulator as sim
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    for case in load_golden("r2_extra.json")["consumption"]:
        eng = _SimFakeEngine()
        pulled = [0]

        def source():
            for i in pyoracle.piece_sequence(case["seed"], case["total"]):
                pulled[0] += 1
                yield PIECES[i]
        src = sim.PieceStream(source()) if mode == "stream" else source()
        policy = sim.policy_best(sim.linear_scorer(collections.OrderedDict((getattr(F, n), w) for n, w in case["weights"])), sim.tie_first)
        kw = {"chunk": 16} if mode == "stream" else {"chunk": 1}
        gen = sim.simulate(State(None, Board(20, 10)), src, sim.move_drop, policy, lookahead=case["lookahead"], engine=eng, **kw)
        moves = 0
        for _ in gen:
            moves += 1
            if case["stop_after"] is not None and moves == case["stop_after"]:
                break
        gen.close()
        assert moves == case["moves"], case
        if mode == "chunk1":
            assert pulled[0] == case["consumed"], (case, pulled[0])
        gen2 = sim.simulate(State(None, Board(20, 10)), src, sim.move_drop, policy, lookahead=case["lookahead"], engine=eng, **kw)
        first = next(gen2, None)
        gen2.close()
        assert (None if first is None else first.delta.zoid.name) == case["second_game_first_piece"], case


def test_state_score_custom_points_matches_reference():
    """State.score(points) (game.py:159-164) of the host State replay, custom tables incl. points[0] != 0."""
    from conftest import load_golden
    from oracle import pyoracle
    from tetris_ai_b200.tetris.game import Board, State, PIECES
    games = load_golden("games.json")["games"]
    for rec in load_golden("r2_extra.json")["points"]:
        g = games[rec["game"]]
        ids = pyoracle.piece_sequence(g["seed"], g["T"])
        tr = bytes.fromhex(g["trace_hex"])
        st = State(None, Board(20, 10))
        for t in range(g["pieces"]):
            st = st.future(PIECES[ids[t]], tr[4 * t], tr[4 * t + 1], tr[4 * t + 2])
        for pts, want in rec["scores"]:
            assert st.score(points=tuple(pts)) == want


def test_host_board_other_sizes_and_device_refusal():
    """game.py:38-45 takes any size; the host model follows, the device boundary (rows16) refuses anything but 20 x 10."""
    from tetris_ai_b200.tetris.game import Board, State, zoids
    b = Board(8, 6)
    assert b.rows() == 8 and b.cols() == 6 and b.data.shape == (8, 6)
    st = State(None, b).future(zoids.classic.O, 0, 6, 0).future(zoids.classic.O, 0, 6, 2).future(zoids.classic.O, 0, 6, 4)
    assert st.lines_cleared() == 2 and st.cleared[2] == 1 and st.board.masks == [0] * 8
    st2 = pickle.loads(pickle.dumps(st))
    assert st2.board.rows() == 8 and st2.board.masks == st.board.masks and st2.score() == st.score() == 100
    with pytest.raises(ValueError):
        b.rows16()


def test_features_header_matches_tables():
    from oracle import pyoracle
    from tetris_ai_b200.tetris import features as F
    h = open(os.path.join(ROOT, "include", "tetris_b200_features.h")).read()
    ids = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"TB200_F_([A-Z0-9_]+?)\s*=\s*(\d+),", h)}
    assert ids == dict(F.FEATURE_ID) == dict(pyoracle.FEATURE_ID)
    names = re.findall(r'"([a-z0-9_]+)"', h[h.index("TB200_FEATURE_NAMES {"):h.index("TB200_W6_FEATURE_IDS")])
    assert names[:45] == list(F.FEATURE_NAMES)


def test_jit_build_compiles_offline_and_caches(tmp_path, monkeypatch):
    """tb200_jit_build: the library's own kernel sources, specialised for a feature list with NVRTC (sm_100a), no device needed;
    the second call finds the cubin in the disk cache."""
    import ctypes
    from tetris_ai_b200 import _lib
    monkeypatch.setenv("TB200_JIT_CACHE", str(tmp_path))
    L = _lib.load()
    ids = np.array([16, 1, 32, 37, 10], dtype=np.int32)          # pits, max_ht, jaggedness, landing_height, row_trans
    log = ctypes.create_string_buffer(1024)
    kib = L.tb200_jit_build(ids.ctypes.data_as(ctypes.c_void_p), len(ids), log, 1024)
    assert kib > 100 and log.value == b"compiled", (kib, log.value)
    kib2 = L.tb200_jit_build(ids.ctypes.data_as(ctypes.c_void_p), len(ids), log, 1024)
    assert kib2 == kib and log.value == b"cached"
    assert len(list(tmp_path.glob("*.cubin"))) == 1
    bad = np.array([99], dtype=np.int32)
    assert L.tb200_jit_build(bad.ctypes.data_as(ctypes.c_void_p), 1, log, 1024) < 0
