#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the dev container (needs /root/reference, which does not exist on the GPU box); the
JSON files it writes are committed so that every test can run without the reference.

The reference (CogWorks/Tetris-AI) ships no scorer, tie-breaker, piece generator or fitness
function, so "bit-exact against the reference" means bit-exact against the reference code driven
by the canonical harness frozen here:

  * pieces      : r = random.Random(seed); ids = [r.randrange(7) for _ in range(T)];
                  id -> list(zoids.classic)[id]  (field order I,T,L,J,O,Z,S; game.py:24-34)
  * scorer      : acc = 0.0; for f, w in weights.items(): acc += float(f(state, cache)) * w
                  (explicit loop: one rounded multiply + one rounded add per feature, in order)
  * tie-breaker : lambda maxes: maxes[0]   (first in enumeration order, simulator.py:125-149)
  * move_gen    : cogworks.tetris.simulator.move_drop (simulator.py:3-13)
  * fitness     : sum(lines_cleared of the candidate's G games) / G

Usage:  python tests/golden/make_golden.py [--skip-ce] [--jobs N]
"""
import argparse
import collections
import collections.abc
import hashlib
import json
import os
import random
import sys
import time

sys.dont_write_bytecode = True
collections.Mapping = collections.abc.Mapping          # removed in py3.10; feature.py:52, learning.py:9
REF = os.environ.get("TETRIS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import numpy as np                                       # noqa: E402
from cogworks import feature as ref_feature              # noqa: E402
from cogworks import learning as ref_learning            # noqa: E402
from cogworks.tetris import features as ref_features     # noqa: E402
from cogworks.tetris import simulator as ref_sim         # noqa: E402
from cogworks.tetris.game import Board, State, zoids     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

FEATURE_NAMES = [
    "mean_ht", "max_ht", "min_ht", "all_ht", "max_ht_diff", "min_ht_diff", "cleared", "cuml_cleared",
    "tetris", "col_trans", "row_trans", "all_trans", "wells", "cuml_wells", "deep_wells", "max_well",
    "pits", "pit_rows", "lumped_pits", "pit_depth", "mean_pit_depth",
    "cd_1", "cd_2", "cd_3", "cd_4", "cd_5", "cd_6", "cd_7", "cd_8", "cd_9",
    "all_diffs", "max_diffs", "jaggedness", "d_max_ht", "d_all_ht", "d_mean_ht", "d_pits",
    "landing_height", "pattern_div", "nine_filled", "tetris_progress", "full_cells", "weighted_cells",
    "eroded_cells", "cuml_eroded",
]
assert len(FEATURE_NAMES) == 45
FEATS = [getattr(ref_features, n) for n in FEATURE_NAMES]
PIECES = list(zoids.classic)                             # I,T,L,J,O,Z,S
PIECE_ID = {z.name: i for i, z in enumerate(PIECES)}

W6 = collections.OrderedDict([
    ("landing_height", -4.500158825082766), ("eroded_cells", 3.4181268101392694),
    ("row_trans", -3.2178882868487753), ("col_trans", -9.348695305445199),
    ("pits", -7.899265427351652), ("cuml_wells", -3.3855972247263626),
])


def weights_by_name(named):
    return collections.OrderedDict((getattr(ref_features, n), w) for n, w in named.items())


def make_scorer(weights):
    items = list(weights.items())

    def scorer(state):
        cache = {}
        acc = 0.0
        for f, w in items:
            acc += float(f(state, cache)) * w
        return acc
    return scorer


def tie_first(maxes):
    return maxes[0]


def piece_ids(seed, T):
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


def board_rows(board):
    """20 ints, row 0 = top; bit c = column c."""
    return [int(sum(1 << c for c in range(board.cols()) if board[r, c])) for r in range(board.rows())]


def feat_value(v):
    """ints exact; floats as hex so that the fixture is bit-exact."""
    if isinstance(v, (float, np.floating)):
        return float(v).hex()
    return int(v)


def all_features(state):
    vals = ref_feature.evaluate(state, FEATS)
    return [feat_value(vals[f]) for f in FEATS]


def play(named_weights, seed, T, lookahead=1, tie=tie_first, keep_trace=True, slide=False, pressure=None, wiggle=None):
    """Canonical game: returns summary + byte trace [rot,row,col,k]*.
    slide=True: move_gen = move_with_overhang_slide(move_drop) (simulator.py:15-47) instead of move_drop.
    pressure=dict(...): wrap the generator in move_with_pressure(gen, **pressure) (simulator.py:88-120)."""
    move_gen = ref_sim.move_with_overhang_slide(ref_sim.move_drop) if slide else ref_sim.move_drop
    if wiggle == "top":                         # move_wiggle: flood fill from the top of every column (simulator.py:81-85)
        move_gen = ref_sim.move_wiggle
    elif wiggle == "drop":                      # move_with_wiggle(move_drop) (simulator.py:49-77)
        move_gen = ref_sim.move_with_wiggle(ref_sim.move_drop)
    if pressure is not None:
        move_gen = ref_sim.move_with_pressure(move_gen, **pressure)
    weights = weights_by_name(named_weights)
    scorer = make_scorer(weights)
    calls = [0]

    def counting(state):
        calls[0] += 1
        return scorer(state)
    ids = piece_ids(seed, T)
    gen = iter([PIECES[i] for i in ids])
    state = State(None, Board(20, 10))
    trace = bytearray()
    n = 0
    last = state
    for st in ref_sim.simulate(state, gen, move_gen, ref_sim.policy_best(counting, tie), lookahead):
        d = st.delta
        trace += bytes([d.rot, d.row, d.col, len(d.cleared)])
        n += 1
        last = st
    score = last.score() if n else 0
    out = {
        "weights": list(named_weights.items()), "seed": seed, "T": T, "lookahead": lookahead,
        "pieces": n, "lines": int(last.lines_cleared()),
        "hist": [int(last.cleared.get(k, 0)) for k in (1, 2, 3, 4)],
        "scorer_calls": calls[0], "score": int(score),
        "trace_sha16": hashlib.sha256(bytes(trace)).hexdigest()[:16],
        "final_board": board_rows(last.board),
        "move_gen": ("wiggle_" + wiggle) if wiggle else ("overhang_slide" if slide else "drop"),
        "pressure": pressure,
    }
    if keep_trace:
        out["trace_hex"] = bytes(trace).hex()
    return out


# ----------------------------------------------------------------------------- feature fixtures
def gen_feature_states(n_games=24, max_pieces=120, seed=2024):
    """Reachable (prev board, piece, rot, col) -> reference move + 45 features + post board."""
    rng = random.Random(seed)
    out = []
    greedy = make_scorer(weights_by_name(W6))
    for g in range(n_games):
        state = State(None, Board(20, 10))
        mode = g % 3                          # 0 random, 1 greedy w6 (many clears), 2 mixed
        for t in range(max_pieces):
            zoid = PIECES[rng.randrange(7)]
            moves = list(ref_sim.move_drop(state, zoid))
            if not moves:
                out.append({"prev": board_rows(state.board), "piece": PIECE_ID[zoid.name], "n_legal": 0})
                break
            futures = [state.future(zoid, *m) for m in moves]
            # record a sample of the futures (all of them for some steps)
            pick = range(len(moves)) if rng.random() < 0.15 else rng.sample(range(len(moves)), min(2, len(moves)))
            for i in pick:
                f = futures[i]
                out.append({
                    "prev": board_rows(state.board), "piece": PIECE_ID[zoid.name],
                    "rot": int(moves[i][0]), "row": int(moves[i][1]), "col": int(moves[i][2]),
                    "k": len(f.delta.cleared), "n_legal": len(moves),
                    "legal_slots": [[int(m[0]), int(m[2]), int(m[1])] for m in moves] if i == pick[0] else None,
                    "post": board_rows(f.board), "features": all_features(f),
                })
            if mode == 0 or (mode == 2 and rng.random() < 0.5):
                nxt = futures[rng.randrange(len(futures))]
            else:
                scores = [greedy(f) for f in futures]
                nxt = futures[scores.index(max(scores))]
            nxt.prev.prev = None              # prune chain (features look one level back only)
            state = nxt
    return out


def known_answer_state():
    """The hand-checked board of SURVEY.md §8a-F."""
    rows = ["#.........", "#..#......", "##.#......", "#########.", "###.#####.", "#########.", "#.#######.", "#########."]
    b = Board(20, 10)
    for i, line in enumerate(rows):
        for c, ch in enumerate(line):
            if ch == "#":
                b.data[12 + i, c] = True
    for c in range(10):
        nz = np.nonzero(b.data[:, c])[0]
        b.heights[c] = int(20 - nz[0]) if len(nz) else 0
    root = State(None, b)
    res = []
    for name, rot, col in (("I", 1, 9), ("J", 1, 8), ("O", 0, 4), ("T", 2, 6)):
        zoid = getattr(zoids.classic, name)
        mv = [m for m in ref_sim.move_drop(root, zoid) if m[0] == rot and m[2] == col][0]
        f = root.future(zoid, *mv)
        res.append({"prev": board_rows(b), "piece": PIECE_ID[name], "rot": int(mv[0]), "row": int(mv[1]), "col": int(mv[2]),
                    "k": len(f.delta.cleared), "n_legal": len(list(ref_sim.move_drop(root, zoid))),
                    "legal_slots": None, "post": board_rows(f.board), "features": all_features(f)})
    return res


# ----------------------------------------------------------------------------- step fixtures
def gen_steps(seed=7, n_games=6, max_pieces=80):
    """best-move fixtures: (board, piece, next) -> chosen move, score bits, n_legal for L1 and L2."""
    rng = random.Random(seed)
    wsets = [dict(W6), {"max_ht": -1.0, "pits": -1.0}, {"max_ht": 0.0},
             {n: random.Random(7 + i).gauss(0, 1) for i, n in enumerate(FEATURE_NAMES)}]
    w45 = collections.OrderedDict()
    r7 = random.Random(7)
    for n in FEATURE_NAMES:
        w45[n] = r7.gauss(0, 1)
    wsets[3] = w45
    out = []
    for g in range(n_games):
        named = wsets[g % len(wsets)]
        weights = weights_by_name(named)
        scorer = make_scorer(weights)
        policy = ref_sim.policy_best(scorer, tie_first)
        state = State(None, Board(20, 10))
        ids = [rng.randrange(7) for _ in range(max_pieces + 1)]
        for t in range(max_pieces):
            z1, z2 = PIECES[ids[t]], PIECES[ids[t + 1]]
            rec = {"board": board_rows(state.board), "piece": ids[t], "next": ids[t + 1],
                   "weights": list(named.items()),
                   "prev_lines": int(state.lines_cleared())}
            # lookahead 1
            l1 = [state.future(z1, *m) for m in ref_sim.move_drop(state, z1)]
            best1 = policy(l1)
            if best1 is None:
                rec["l1"] = None
                rec["l2"] = None
                out.append(rec)
                break
            rec["l1"] = {"rot": int(best1.delta.rot), "row": int(best1.delta.row), "col": int(best1.delta.col),
                         "k": len(best1.delta.cleared), "score": float(scorer(best1)).hex(), "n_legal": len(l1)}
            # lookahead 2 (one simulate() iteration, simulator.py:156-180)
            if t % 3 == 0:
                l2 = [p.future(z2, *m) for p in l1 for m in ref_sim.move_drop(p, z2)]
                if l2:
                    leaf = policy(l2)
                    first = leaf
                    while first.prev is not state:
                        first = first.prev
                    rec["l2"] = {"rot": int(first.delta.rot), "row": int(first.delta.row), "col": int(first.delta.col),
                                 "k": len(first.delta.cleared), "score": float(scorer(leaf)).hex(),
                                 "n_leaves": len(l2)}
                else:
                    rec["l2"] = {"rot": int(best1.delta.rot), "row": int(best1.delta.row), "col": int(best1.delta.col),
                                 "k": len(best1.delta.cleared), "score": float(scorer(best1)).hex(), "n_leaves": 0}
            else:
                rec["l2"] = None
            out.append(rec)
            # advance with a mix of best and random moves to diversify boards
            nxt = best1 if rng.random() < 0.7 else l1[rng.randrange(len(l1))]
            nxt.prev.prev = None
            state = nxt
    return out


def gen_slide_moves(seed=5, n_games=10, max_pieces=60):
    """Boards with overhangs -> the exact move list of move_with_overhang_slide(move_drop) for every piece."""
    rng = random.Random(seed)
    gen_slide = ref_sim.move_with_overhang_slide(ref_sim.move_drop)
    out = []
    for g in range(n_games):
        state = State(None, Board(20, 10))
        for t in range(max_pieces):
            zoid = PIECES[rng.randrange(7)]
            moves = list(gen_slide(state, zoid))
            if not moves:
                break
            base = list(ref_sim.move_drop(state, zoid))
            if len(moves) > len(base) or rng.random() < 0.1:
                out.append({"board": board_rows(state.board),
                            "moves": [[[int(x) for x in m] for m in gen_slide(state, z)] for z in PIECES]})
            state = state.future(zoid, *moves[rng.randrange(len(moves))])
            state.prev = None
    return out


def gen_wiggle_moves(seed=6, n_games=6, max_pieces=50):
    """Boards (with overhangs / tucks) -> exact move lists of move_wiggle and move_with_wiggle(move_drop) for every piece."""
    rng = random.Random(seed)
    drop_w = ref_sim.move_with_wiggle(ref_sim.move_drop)
    out = []
    for g in range(n_games):
        state = State(None, Board(20, 10))
        for t in range(max_pieces):
            zoid = PIECES[rng.randrange(7)]
            moves = list(ref_sim.move_wiggle(state, zoid))
            if not moves:
                break
            if t % 4 == 0 or len(moves) > len(list(ref_sim.move_drop(state, zoid))):
                out.append({"board": board_rows(state.board),
                            "top": [[[int(x) for x in m] for m in ref_sim.move_wiggle(state, z)] for z in PIECES],
                            "drop": [[[int(x) for x in m] for m in drop_w(state, z)] for z in PIECES]})
            state = state.future(zoid, *moves[rng.randrange(len(moves))])
            state.prev = None
    return out


# ----------------------------------------------------------------------------- CE fixture
_CE_FN = None


def _ce_worker(i):
    return _CE_FN(i)


def fork_map(jobs):
    import multiprocessing as mp

    def map_f(fn, it):
        global _CE_FN
        _CE_FN = fn
        with mp.get_context("fork").Pool(jobs) as pool:
            return pool.map(_ce_worker, list(it), chunksize=1)
    return map_f


def ce_test_f(G=10, seed0=1000, T=1000, log=None):
    seqs = [piece_ids(seed0 + g, T) for g in range(G)]

    def test_f(weights):
        scorer = make_scorer(weights)
        policy = ref_sim.policy_best(scorer, tie_first)
        lines, pieces = [], []
        for ids in seqs:
            last = State(None, Board(20, 10))
            n = 0
            for st in ref_sim.simulate(last, iter([PIECES[i] for i in ids]), ref_sim.move_drop, policy, 1):
                st.prev.prev = None
                last = st
                n += 1
            lines.append(int(last.lines_cleared()))
            pieces.append(n)
        return sum(lines) / G
    return test_f


def gen_ce(jobs, generations=2):
    feats = [getattr(ref_features, n) for n in W6.keys()]
    rng = random.Random(42)
    fitness_log = []
    base_map = fork_map(jobs)

    def map_f(fn, it):
        res = base_map(fn, it)
        fitness_log.append([r[0] for r in res])
        return res
    gen = ref_learning.cross_entropy(feats, 10.0, 100, 10, rng, ce_test_f(), map_f=map_f)
    out = {"features": list(W6.keys()), "stdev": 10.0, "width": 100, "keep": 10, "rng_seed": 42,
           "games": 10, "seed0": 1000, "T": 1000, "lookahead": 1, "generations": []}
    for g in range(generations):
        t0 = time.time()
        mean, std = next(gen)
        out["generations"].append({
            "fitness": [float(x).hex() for x in fitness_log[g]],
            "mean": [float(v).hex() for v in mean.values()],
            "std": [float(v).hex() for v in std.values()],
            "mean_repr": [repr(float(v)) for v in mean.values()],
            "std_repr": [repr(float(v)) for v in std.values()],
            "wall_s": round(time.time() - t0, 1),
        })
        print("CE generation", g, "done in", round(time.time() - t0, 1), "s", flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-ce", action="store_true")
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    args = ap.parse_args()

    def dump(name, obj):
        import gzip
        data = (json.dumps(obj, separators=(",", ":")) + "\n").encode()
        if name.endswith(".gz"):
            with open(os.path.join(HERE, name), "wb") as raw:      # mtime=0: reproducible bytes
                with gzip.GzipFile(fileobj=raw, mode="wb", compresslevel=9, mtime=0) as f:
                    f.write(data)
        else:
            with open(os.path.join(HERE, name), "wb") as f:
                f.write(data)
        print("wrote", name, os.path.getsize(os.path.join(HERE, name)), "bytes", flush=True)

    meta = {"reference": "CogWorks/Tetris-AI (cogworks 2.0a2) run from " + REF,
            "python": sys.version.split()[0], "numpy": np.__version__,
            "feature_names": FEATURE_NAMES, "piece_order": [z.name for z in PIECES]}

    t0 = time.time()
    states = known_answer_state() + gen_feature_states()
    dump("feature_states.json.gz", {"meta": meta, "states": states})
    print("feature states:", len(states), round(time.time() - t0, 1), "s", flush=True)

    t0 = time.time()
    steps = gen_steps()
    dump("steps.json", {"meta": meta, "steps": steps})
    print("steps:", len(steps), round(time.time() - t0, 1), "s", flush=True)

    t0 = time.time()
    weak = collections.OrderedDict([("max_ht", -1.0), ("pits", -1.0)])
    ties = collections.OrderedDict([("max_ht", 0.0)])
    r7 = random.Random(7)
    w45 = collections.OrderedDict((n, r7.gauss(0, 1)) for n in FEATURE_NAMES)
    games = []
    for s in (0, 1, 2):
        games.append(play(W6, s, 500))
    games.append(dict(play(W6, 0, 500, tie=lambda m: m[-1]), tie="last"))
    games.append(play(W6, 0, 60, lookahead=2))
    for s in (0, 1, 2):
        games.append(play(weak, s, 5000))
    games.append(play(ties, 0, 5000))
    games.append(play(ties, 0, 5000, lookahead=2))
    games.append(play(weak, 3, 5000, lookahead=2))
    for s in (0, 1):
        games.append(play(w45, s, 200))
    for s in (0, 1):
        games.append(play(w45, s, 40, lookahead=2))
    # full feature set, but dominated by the w6 weights so that games are long and clear lines
    r8 = random.Random(8)
    w45b = collections.OrderedDict((n, W6.get(n, 0.0) + r8.gauss(0, 0.3)) for n in FEATURE_NAMES)
    for s in (0, 1):
        games.append(play(w45b, s, 300))
    games.append(play(w45b, 0, 30, lookahead=2))
    dump("games.json", {"meta": meta, "games": games})

    # §8f rank 2: move_with_overhang_slide(move_drop).  Weight sets that leave overhangs (so that slides happen).
    t0 = time.time()
    holes = collections.OrderedDict([("max_ht", -1.0), ("row_trans", -0.2), ("pits", 0.4), ("landing_height", -0.5)])
    sgames = []
    for s in (0, 1, 2):
        sgames.append(play(W6, s, 300, slide=True))
    for s in (0, 1, 2, 3):
        sgames.append(play(holes, s, 400, slide=True))
    sgames.append(play(weak, 1, 5000, slide=True))
    sgames.append(play(holes, 0, 40, lookahead=2, slide=True))
    sgames.append(play(W6, 1, 40, lookahead=2, slide=True))
    sgames.append(play(w45b, 1, 200, slide=True))
    dump("games_slide.json", {"meta": meta, "games": sgames})
    print("slide games:", len(sgames), round(time.time() - t0, 1), "s", flush=True)
    dump("slide_moves.json.gz", {"meta": meta, "boards": gen_slide_moves()})

    # §8f rank 3: move_with_pressure (human time-pressure filter, level-dependent)
    t0 = time.time()
    pgames = []
    for s in (0, 1):
        pgames.append(play(W6, s, 1500, pressure={}))                     # default parameters; ends when the level gets too fast
    pgames.append(play(W6, 2, 1500, pressure={"two_but_rot": True}))
    pgames.append(play(W6, 3, 1500, pressure={"avg_lat": 180.0, "resp_time": 60.5, "move_eff": 1.1}))
    pgames.append(play(W6, 4, 1500, slide=True, pressure={}))
    pgames.append(play(holes, 0, 400, slide=True, pressure={"two_but_rot": True}))
    pgames.append(play(W6, 0, 60, lookahead=2, pressure={}))
    pgames.append(play(w45b, 0, 400, pressure={"avg_lat": 300.0}))
    dump("games_pressure.json", {"meta": meta, "games": pgames})

    # §8f rank 4: move_wiggle / move_with_wiggle(move_drop) (DFS flood fill; the move ORDER is the recursion order)
    t0 = time.time()
    wgames = []
    for s in (0, 1):
        wgames.append(play(W6, s, 120, wiggle="top"))
    for s in (0, 1, 2):
        wgames.append(play(holes, s, 300, wiggle="top"))
    wgames.append(play(holes, 3, 300, wiggle="drop"))
    wgames.append(play(W6, 2, 120, wiggle="drop"))
    wgames.append(play(ties, 0, 200, wiggle="top"))
    wgames.append(play(W6, 3, 150, wiggle="top", pressure={}))
    wgames.append(play(holes, 0, 25, lookahead=2, wiggle="top"))
    wgames.append(play(W6, 1, 20, lookahead=2, wiggle="drop"))
    wgames.append(play(holes, 2, 25, lookahead=2, wiggle="top", pressure={"two_but_rot": True}))
    dump("games_wiggle.json", {"meta": meta, "games": wgames})
    print("wiggle games:", len(wgames), round(time.time() - t0, 1), "s", flush=True)
    dump("wiggle_moves.json.gz", {"meta": meta, "boards": gen_wiggle_moves()})
    print("pressure games:", len(pgames), round(time.time() - t0, 1), "s", flush=True)
    print("games:", len(games), round(time.time() - t0, 1), "s", flush=True)

    if not args.skip_ce:
        dump("ce.json", {"meta": meta, "ce": gen_ce(args.jobs)})


if __name__ == "__main__":
    main()
