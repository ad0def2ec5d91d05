"""GPU-backed mirror of cogworks/tetris/simulator.py for the hard-drop path.

  move_drop(state, zoid)                       simulator.py:3    legal (rot,row,col) in rot-major, col-minor order
  policy_best(scorer, tie_breaker)             simulator.py:123  argmax with the first-of-max tie rule
  simulate(state, zoid_gen, move_gen, move_policy, lookahead=1)   simulator.py:152  generator of States

The search (enumerate, drop, clear, features, fp64 score, argmax, commit) runs in the K2 kernel;
this module only adapts the reference's call signatures.  Only the combination the hot path is
defined for is accepted — move_gen = move_drop and a policy built by policy_best(linear_scorer(...),
tie_first); anything else raises, because this engine has no CPU fallback.
"""
from __future__ import annotations

import collections
import collections.abc
import itertools

import numpy as np

from ..feature import default_engine
from .features import Feature
from .game import PIECES, State

CHUNK = 4096          # pieces handed to the device per launch by simulate()


def move_drop(state, zoid, engine=None):
    """Yield every legal hard-drop placement (rot, row, col) of `zoid` on state.board, in the reference's order."""
    eng = engine or default_engine()
    slots = [(rot, col) for rot in range(len(zoid)) for col in range(state.board.cols() - zoid[rot].shape[1] + 1)]
    n = len(slots)
    r = eng.eval_features(np.tile(state.board.rows16(), (n, 1)), [zoid.id] * n, [s[0] for s in slots], [s[1] for s in slots])
    for (rot, col), legal, row in zip(slots, r["legal"], r["row"]):
        if legal == 1:
            yield (rot, int(row), col)


class OverhangSlide:
    """move_with_overhang_slide(move_drop) (simulator.py:15-47) as a device move set: every hard drop followed by the
    placements reached by sliding under an overhang at the landed row.  It is consumed by simulate(); enumerating the
    moves on the host would need a CPU search, which this engine does not have."""
    move_set = 1

    def __call__(self, state, zoid):
        raise TypeError("the overhang-slide move set is evaluated on the device; pass it to simulate() as move_gen")


def move_with_overhang_slide(move_gen):
    if move_gen is not move_drop:
        raise TypeError("only move_with_overhang_slide(move_drop) is implemented on the device (no CPU fallback)")
    return OverhangSlide()


class Wiggle:
    """move_with_wiggle(move_drop) / move_wiggle (simulator.py:49-85) as device move sets: every position the piece can
    reach by moving down, left and right, in the order of the reference's recursion."""

    def __init__(self, move_set):
        self.move_set = move_set

    def __call__(self, state, zoid):
        raise TypeError("the wiggle move sets are evaluated on the device; pass them to simulate() as move_gen")


def move_with_wiggle(move_gen):
    if move_gen is not move_drop:
        raise TypeError("only move_with_wiggle(move_drop) and move_wiggle are implemented on the device (no CPU fallback)")
    return Wiggle(8)


move_wiggle = Wiggle(4)


class Pressure:
    """move_with_pressure(move_gen, avg_lat, resp_time, move_eff, two_but_rot) (simulator.py:88-120) as a device move set:
    the level-dependent time-pressure filter, applied to the moves of move_drop or of move_with_overhang_slide(move_drop)."""

    def __init__(self, inner_move_set, avg_lat, resp_time, move_eff, two_but_rot):
        self.move_set = inner_move_set
        self.params = {"avg_lat": avg_lat, "resp_time": resp_time, "move_eff": move_eff, "two_but_rot": two_but_rot}

    def __call__(self, state, zoid):
        raise TypeError("the pressure move set is evaluated on the device; pass it to simulate() as move_gen")


def move_with_pressure(move_gen, avg_lat=250.9, resp_time=79.8, move_eff=1.23, two_but_rot=False):
    if move_gen is move_drop:
        inner = 0
    elif isinstance(move_gen, (OverhangSlide, Wiggle)):
        inner = move_gen.move_set
    else:
        raise TypeError("move_with_pressure is implemented on the device around move_drop, move_with_overhang_slide(move_drop), "
                        "move_with_wiggle(move_drop) and move_wiggle only")
    return Pressure(inner, avg_lat, resp_time, move_eff, two_but_rot)


class LinearScorer:
    """The canonical scorer: acc = 0.0; for f, w in weights.items(): acc += float(f(state)) * w (device fp64, no FMA)."""

    def __init__(self, weights):
        if not isinstance(weights, collections.abc.Mapping) or not weights:
            raise TypeError("linear_scorer needs a non-empty Mapping feature -> weight")
        for f in weights:
            if not isinstance(f, Feature):
                raise TypeError("feature %r is not a device feature; the B200 engine has no CPU fallback" % (f,))
        self.features = list(weights.keys())
        self.weights = [float(w) for w in weights.values()]

    def __call__(self, state):
        from ..feature import evaluate
        acc = 0.0
        for v in evaluate(state, dict(zip(self.features, self.weights))).values():
            acc += v
        return acc


def linear_scorer(weights):
    return LinearScorer(weights)


def tie_first(maxes):
    """Canonical tie-breaker: first element of the max group (lowest rot, then lowest col)."""
    return maxes[0]


class BestPolicy:
    def __init__(self, scorer, tie_breaker):
        if not isinstance(scorer, LinearScorer):
            raise TypeError("policy_best on the B200 engine needs linear_scorer(weights); arbitrary Python scorers "
                            "would need a CPU path, which this engine does not have")
        if tie_breaker is not tie_first:
            raise TypeError("only tie_first (first of the max group in enumeration order) is implemented on the device")
        self.scorer = scorer

    def __call__(self, futures):
        raise TypeError("a device policy is consumed by simulate(); it does not score host-side State lists")


def policy_best(scorer, tie_breaker=tie_first):
    return BestPolicy(scorer, tie_breaker)


class PieceStream:
    """Iterator with push-back for piece generators that are SHARED between successive simulate() calls.

    The reference's simulate() pulls exactly `lookahead` pieces from zoid_gen per move (simulator.py:156).  This engine
    hands the device a whole chunk of pieces per launch, so it reads ahead; when zoid_gen is a PieceStream, simulate()
    returns every piece beyond the reference's own consumption point to it on exit (normal end, game over, or the caller
    closing the generator), so the next game sees exactly the pieces it would see with the reference.  Plain iterators
    cannot take pieces back: pass chunk=1 to simulate() for reference-identical consumption, or accept the read-ahead.
    """

    def __init__(self, iterable):
        self._it = iter(iterable)
        self._front = collections.deque()

    def __iter__(self):
        return self

    def __next__(self):
        if self._front:
            return self._front.popleft()
        return next(self._it)

    def push_back(self, pieces):
        self._front.extendleft(reversed(list(pieces)))


def simulate(state, zoid_gen, move_gen, move_policy, lookahead=1, engine=None, chunk=None):
    """Generator of successive States, identical to the reference's for the accepted arguments.

    chunk: pieces handed to the device per launch (default CHUNK).  Read-ahead (documented deviation): up to
    `chunk` pieces are taken from zoid_gen before each launch, where the reference takes `lookahead` per move.  The
    yielded States never depend on it.  To keep a shared generator in step with the reference, wrap it in PieceStream
    (unconsumed pieces are pushed back) or pass chunk=1 (one launch per move, reference-identical consumption).
    """
    if lookahead not in (1, 2):
        raise ValueError("lookahead must be 1 or 2 on the B200 engine (TB200_MAX_LOOKAHEAD; the reference accepts any lookahead > 0, "
                         "simulator.py:153, its configs use 1 and 2)")
    pressure = None
    if move_gen is move_drop:
        move_set = 0
    elif isinstance(move_gen, (OverhangSlide, Wiggle)):
        move_set = move_gen.move_set
    elif isinstance(move_gen, Pressure):
        move_set, pressure = move_gen.move_set, move_gen.params
    else:
        raise TypeError("only move_drop and move_with_overhang_slide(move_drop) are implemented on the device (no CPU fallback)")
    if not isinstance(move_policy, BestPolicy):
        raise TypeError("move_policy must come from policy_best(linear_scorer(weights), tie_first)")
    eng = engine or default_engine()
    sc = move_policy.scorer
    size = max(int(chunk if chunk is not None else CHUNK), 1) + (lookahead - 1)      # pieces per launch incl. the look-ahead piece
    source = zoid_gen if hasattr(zoid_gen, "push_back") else None
    zoid_gen = iter(zoid_gen)
    window = []              # pieces read from zoid_gen and not yet placed
    unread = 0               # how many pieces at the tail of `window` the reference would not have pulled yet
    try:
        while True:
            window = window + list(itertools.islice(zoid_gen, 0, size - len(window)))
            if not window:
                return
            more = len(window) == size
            # with lookahead 2 the last piece of a non-final chunk is only looked at, not placed
            commit = len(window) - 1 if (more and lookahead == 2) else len(window)
            unread = len(window)
            ids = np.array([[z.id for z in window]], dtype=np.uint8)
            eng.set_features(sc.features)
            r = eng.play_games(np.array([sc.weights]), 1, pieces=ids, lookahead=lookahead, max_moves=commit,
                               init_rows=[state.board.rows16()], want_trace=True, want_fitness=False, move_set=move_set, pressure=pressure,
                               init_lines=state.lines_cleared())
            placed = int(r["results"]["pieces"][0])
            for t in range(placed):
                rot, row, col, _k = (int(x) for x in r["trace"][0, t])
                state = state.future(window[t], rot, row, col)
                # the reference has pulled t + lookahead pieces of this window when it yields move t (simulator.py:156,184)
                unread = max(len(window) - (t + lookahead), 0)
                yield state
            if placed < commit:
                unread = max(len(window) - (placed + lookahead), 0)      # game over: it pulled the `lookahead` visible pieces and stopped
                return
            if not more:
                unread = 0
                return
            window = window[commit:]
            unread = 0
    finally:
        if source is not None and unread:
            source.push_back(window[len(window) - unread:])


__all__ = ["PieceStream", "move_drop", "move_with_overhang_slide", "move_with_pressure", "move_with_wiggle", "move_wiggle", "policy_best", "simulate", "linear_scorer", "tie_first", "LinearScorer", "BestPolicy", "PIECES"]
