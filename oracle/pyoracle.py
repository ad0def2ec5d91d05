"""pyoracle — CPU restatement (numpy) of the CogWorks/Tetris-AI placement hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under tetris_ai_b200/ imports this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it, and there
only as the checker or the timed CPU baseline — never as the product path.

Parity status: PINNED BY REFERENCE OUTPUTS.  The reference ships no tests or golden vectors of its
own, so this restatement is pinned against outputs of the unmodified reference
generated in the dev container by tests/golden/make_golden.py (feature vectors of 7k reachable
states, best-move steps, whole-game traces, two cross-entropy generations) — see
tests/test_oracle_golden.py.

It keeps the reference's data representation on purpose (a rows x cols numpy bool grid with row 0
at the TOP plus a cached per-column height list) so that its speed is representative of the
reference's own CPU path; the CUDA kernels use a completely different (column-bitmask) layout.
Every function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import collections
import heapq
import itertools
import random

import numpy as np

ROWS, COLS = 20, 10

# ---------------------------------------------------------------------------------------------
# Pieces.  cogworks/tetris/game.py:6-34: rotation i of a piece is np.rot90(shape, k=-i)
# (clockwise); namedtuple field order (= piece id order) is I,T,L,J,O,Z,S with 2,4,4,4,1,2,2
# rotations.
# ---------------------------------------------------------------------------------------------
PIECE_NAMES = ("I", "T", "L", "J", "O", "Z", "S")
_BASE = {
    "I": (["####"], 2),
    "T": (["###", ".#."], 4),
    "L": (["###", "#.."], 4),
    "J": (["###", "..#"], 4),
    "O": (["##", "##"], 1),
    "Z": (["##.", ".##"], 2),
    "S": ([".##", "##."], 2),
}


def _cw(a):
    """One clockwise quarter turn: new[i][j] = old[h-1-j][i]  (== np.rot90(a, k=-1))."""
    h, w = a.shape
    out = np.zeros((w, h), dtype=np.bool_)
    for i in range(w):
        for j in range(h):
            out[i, j] = a[h - 1 - j, i]
    return out


def _build_pieces():
    pieces = []
    for name in PIECE_NAMES:
        rows, nrot = _BASE[name]
        a = np.array([[ch == "#" for ch in r] for r in rows], dtype=np.bool_)
        rots = []
        for _ in range(nrot):
            rots.append(a)
            a = _cw(a)
        pieces.append(tuple(rots))
    return tuple(pieces)


PIECES = _build_pieces()          # PIECES[piece_id][rot] -> bool array (h, w)


# ---------------------------------------------------------------------------------------------
# Board.  cogworks/tetris/game.py:38-108
# ---------------------------------------------------------------------------------------------
class Grid:
    __slots__ = ("cells", "tops")

    def __init__(self, cells=None, tops=None):
        self.cells = np.zeros((ROWS, COLS), dtype=np.bool_) if cells is None else cells
        self.tops = [0] * COLS if tops is None else tops      # column heights (game.py:43)

    def clone(self):                                          # game.py:104-108
        return Grid(self.cells.copy(), list(self.tops))

    @staticmethod
    def from_rows(rows):
        """rows: 20 ints, row 0 = top, bit c = column c."""
        g = Grid()
        for r, bits in enumerate(rows):
            for c in range(COLS):
                g.cells[r, c] = bool((bits >> c) & 1)
        g.recompute_tops()
        return g

    def to_rows(self):
        return [int(sum(1 << c for c in range(COLS) if self.cells[r, c])) for r in range(ROWS)]

    def recompute_tops(self):                                 # game.py:82-87
        for c in range(COLS):
            nz = np.nonzero(self.cells[:, c])[0]
            self.tops[c] = ROWS - nz[0] if len(nz) > 0 else 0

    def collides(self, shape, row, col):                      # game.py:69-75 (overlaps)
        h, w = shape.shape
        if row < 0 or col < 0 or row + h > ROWS or col + w > COLS:
            return True
        return bool(np.any(self.cells[row:row + h, col:col + w] & shape))

    def stamp(self, shape, row, col):                         # game.py:62-67 (imprint)
        h, w = shape.shape
        self.cells[row:row + h, col:col + w] |= shape
        for j in range(w):
            top = ROWS - row - np.nonzero(shape[:, j])[0][0]
            if top >= self.tops[col + j]:
                self.tops[col + j] = top

    def full_rows(self):                                      # game.py:77-78
        return frozenset(int(r) for r in np.nonzero(np.all(self.cells, axis=1))[0])

    def remove_rows(self, rows):                              # game.py:80-87 (clear)
        keep = np.delete(self.cells, sorted(rows), axis=0)
        self.cells = np.concatenate([np.zeros((len(rows), COLS), dtype=np.bool_), keep], axis=0)
        self.recompute_tops()


# ---------------------------------------------------------------------------------------------
# State.  cogworks/tetris/game.py:110-180
# ---------------------------------------------------------------------------------------------
class Node:
    __slots__ = ("prev", "grid", "hist", "piece", "rot", "row", "col", "gone")

    def __init__(self, prev, grid, hist=None, piece=None, rot=None, row=None, col=None, gone=None):
        self.prev = prev
        self.grid = grid
        self.hist = hist if hist is not None else {}          # {k simultaneous rows: times}
        self.piece, self.rot, self.row, self.col = piece, rot, row, col
        self.gone = gone                                      # frozenset of pre-clear full rows

    def lines(self):                                          # game.py:150-154
        return sum(k * n for k, n in self.hist.items())

    def level(self):                                          # game.py:156-157
        return self.lines() // 10

    def child(self, piece, rot, row, col):                    # game.py:166-180 (future)
        g = self.grid.clone()
        hist = dict(self.hist)
        g.stamp(PIECES[piece][rot], row, col)
        full = g.full_rows()
        if full:
            hist[len(full)] = hist.get(len(full), 0) + 1
            g.remove_rows(full)
        return Node(self, g, hist, piece, rot, row, col, full)


def game_score(node, points=(0, 40, 100, 300, 1200)):         # game.py:159-164
    total = 0
    while node.piece is not None:
        total += points[len(node.gone)] * (node.level() + 1)
        node = node.prev
    return total


# ---------------------------------------------------------------------------------------------
# Move generator.  cogworks/tetris/simulator.py:3-13 (move_drop): rot-major, col-minor; legal iff
# the hover row above the stack is in bounds; then lower while free.
# ---------------------------------------------------------------------------------------------
def drop_moves(node, piece):
    g = node.grid
    out = []
    for rot, shape in enumerate(PIECES[piece]):
        h, w = shape.shape
        for col in range(COLS - w + 1):
            row = ROWS - max(g.tops[col:col + w]) - h
            if not g.collides(shape, row, col):
                while not g.collides(shape, row + 1, col):
                    row += 1
                out.append((rot, int(row), col))
    return out


def slide_moves(node, piece):
    """cogworks/tetris/simulator.py:15-47 move_with_overhang_slide(move_drop): each hard-drop move, then the
    placements reached by sliding sideways at the landed row under an overhang — right slides in ascending
    column order (:25-33), then left slides in descending column order (:36-45) — each lowered as far as it
    goes; a direction stops at the first blocked column."""
    g = node.grid
    out = []
    for rot, row, col in drop_moves(node, piece):
        out.append((rot, row, col))
        shape = PIECES[piece][rot]
        w = shape.shape[1]
        if col + w < COLS and row > ROWS - g.tops[col + w]:
            for c in range(col + 1, COLS - w + 1):
                if g.collides(shape, row, c):
                    break
                r = row
                while not g.collides(shape, r + 1, c):
                    r += 1
                out.append((rot, int(r), c))
        if col > 0 and row > ROWS - g.tops[col - 1]:
            for c in range(col - 1, -1, -1):
                if g.collides(shape, row, c):
                    break
                r = row
                while not g.collides(shape, r + 1, c):
                    r += 1
                out.append((rot, int(r), c))
    return out


def _wiggle(node, piece, starts):
    """cogworks/tetris/simulator.py:49-77 move_with_wiggle: depth-first flood fill (down, then left, then right) from each
    start with a `seen` set shared by all starts; a position that cannot move down is a move; the list order is the
    recursion order."""
    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 5000))
    g = node.grid
    seen = set()
    out = []

    def rec(rot, row, col):
        if (rot, row, col) in seen:
            return
        seen.add((rot, row, col))
        shape = PIECES[piece][rot]
        if g.collides(shape, row, col):
            return
        if g.collides(shape, row + 1, col):
            out.append((rot, row, col))
        else:
            rec(rot, row + 1, col)
        if col > 0:
            rec(rot, row, col - 1)
        if col + shape.shape[1] <= COLS:
            rec(rot, row, col + 1)
    for rot, row, col in starts:
        rec(rot, row, col)
    return out


def wiggle_moves_top(node, piece):        # simulator.py:81-85 move_wiggle
    return _wiggle(node, piece, [(rot, 0, col) for rot, shape in enumerate(PIECES[piece]) for col in range(COLS - shape.shape[1] + 1)])


def wiggle_moves_drop(node, piece):       # move_with_wiggle(move_drop)
    return _wiggle(node, piece, drop_moves(node, piece))


MOVE_GENERATORS = {"drop": drop_moves, "slide": slide_moves, "wiggle_top": wiggle_moves_top, "wiggle_drop": wiggle_moves_drop}

_START_COL = {"O": [4], "L": [4, 4, 4, 5], "J": [4, 4, 4, 5], "S": [4, 5], "Z": [4, 5], "T": [4, 4, 4, 5], "I": [3, 5]}
_BRACKETS = [(29, 20), (19, 30), (16, 50), (13, 70), (10, 80), (9, 100), (8, 130), (7, 220), (6, 300), (5, 380),
             (4, 470), (3, 550), (2, 630), (1, 720), (0, 800)]


def pressure_filter(move_gen, avg_lat=250.9, resp_time=79.8, move_eff=1.23, two_but_rot=False):
    """cogworks/tetris/simulator.py:88-120 move_with_pressure: keep the moves that can be keyed in before the piece
    locks at the speed of the CURRENT level (node.level() of the state the moves are generated from)."""
    def filtered(node, piece):
        lvl = node.level()
        speed = next(sp for (lv, sp) in _BRACKETS if lv <= lvl)                      # :113-114
        out = []
        for rot, row, col in move_gen(node, piece):
            clicks = abs(_START_COL[PIECE_NAMES[piece]][rot] - col) + (1 if two_but_rot and rot == 3 else rot)   # :104
            t = resp_time + (clicks * move_eff * avg_lat)                            # :107
            if t <= (ROWS - row) * speed:                                            # :116-118
                out.append((rot, row, col))
        return out
    return filtered


# ---------------------------------------------------------------------------------------------
# Features.  cogworks/tetris/features.py:18-347; ids = source order.
# ---------------------------------------------------------------------------------------------
FEATURE_NAMES = (
    "mean_ht", "max_ht", "min_ht", "all_ht", "max_ht_diff", "min_ht_diff", "cleared", "cuml_cleared",
    "tetris", "col_trans", "row_trans", "all_trans", "wells", "cuml_wells", "deep_wells", "max_well",
    "pits", "pit_rows", "lumped_pits", "pit_depth", "mean_pit_depth",
    "cd_1", "cd_2", "cd_3", "cd_4", "cd_5", "cd_6", "cd_7", "cd_8", "cd_9",
    "all_diffs", "max_diffs", "jaggedness", "d_max_ht", "d_all_ht", "d_mean_ht", "d_pits",
    "landing_height", "pattern_div", "nine_filled", "tetris_progress", "full_cells", "weighted_cells",
    "eroded_cells", "cuml_eroded",
)
FEATURE_ID = {n: i for i, n in enumerate(FEATURE_NAMES)}
NUM_FEATURES = len(FEATURE_NAMES)


def _tops(node):                      # features.py:18-20
    return [int(t) for t in node.grid.tops]


def _wells(node):                     # features.py:106-112
    t = [ROWS] + _tops(node) + [ROWS]
    return [max(0, min(t[i - 1], t[i + 1]) - t[i]) for i in range(1, COLS + 1)]


def _pits(node):                      # features.py:146-154  (rows strictly below each column top)
    t = _tops(node)
    cells = node.grid.cells
    return [{r for r in range(ROWS - t[c] + 1, ROWS) if not cells[r, c]} for c in range(COLS)]


def _nine(node):                      # features.py:279-286  row -> its single empty column
    out = {}
    cells = node.grid.cells
    for r in range(ROWS):
        empty = np.nonzero(~cells[r])[0]
        if len(empty) == 1:
            out[r] = int(empty[0])
    return out


def _mean_ht(node):
    t = _tops(node)
    return sum(t) / len(t)            # features.py:24-26 (true division)


def _pit_total(node):
    return sum(len(s) for s in _pits(node))


def _pit_depth(node):                 # features.py:182-188
    t = _tops(node)
    p = _pits(node)
    return sum(len(set(range(ROWS - t[c], r)) - p[c]) for c in range(COLS) for r in p[c])


def _tetris_progress(node):           # features.py:297-319
    nine = _nine(node)
    cells = node.grid.cells
    progress = 0
    covered = set()
    for r in range(ROWS):
        covered |= {int(c) for c in np.nonzero(cells[r])[0]}
        hole = nine.get(r)
        if hole is not None:
            if hole not in covered:
                progress += 1
            else:
                break
        elif len(covered) < COLS:
            progress = 0
        else:
            break
    return progress


def _eroded(node):                    # features.py:335-340
    shape = PIECES[node.piece][node.rot]
    return int(sum(int(shape[r - node.row].sum()) for r in node.gone))


def feature(node, fid):
    """Value of feature `fid` on `node`, with the reference's int/float typing."""
    n = FEATURE_NAMES[fid]
    if n == "mean_ht":
        return _mean_ht(node)
    if n == "max_ht":
        return max(_tops(node))
    if n == "min_ht":
        return min(_tops(node))
    if n == "all_ht":
        return sum(_tops(node))
    if n == "max_ht_diff":            # features.py:48-50
        return max(_tops(node)) - _mean_ht(node)
    if n == "min_ht_diff":            # features.py:54-56
        return _mean_ht(node) - min(_tops(node))
    if n == "cleared":                # features.py:62-64
        return len(node.gone)
    if n == "cuml_cleared":           # features.py:69-71
        k = len(node.gone)
        return k * (k + 1) / 2
    if n == "tetris":                 # features.py:75-77
        return len(node.gone) // 4
    if n == "col_trans":              # features.py:83-85  (bool diff == xor, no floor/ceiling)
        c = node.grid.cells
        return int(np.sum(c[1:] ^ c[:-1]))
    if n == "row_trans":              # features.py:89-91  (no walls)
        c = node.grid.cells
        return int(np.sum(c[:, 1:] ^ c[:, :-1]))
    if n == "all_trans":
        return feature(node, FEATURE_ID["col_trans"]) + feature(node, FEATURE_ID["row_trans"])
    if n == "wells":
        return sum(_wells(node))
    if n == "cuml_wells":             # features.py:123-125
        return sum(x * (x + 1) / 2 for x in _wells(node))
    if n == "deep_wells":
        return sum(x for x in _wells(node) if x > 1)
    if n == "max_well":
        return max(_wells(node))
    if n == "pits":
        return _pit_total(node)
    if n == "pit_rows":               # features.py:164-166
        return len({r for s in _pits(node) for r in s})
    if n == "lumped_pits":            # features.py:171-178
        return sum(1 for s in _pits(node) for r in s if r - 1 not in s)
    if n == "pit_depth":
        return _pit_depth(node)
    if n == "mean_pit_depth":         # features.py:192-194
        p = _pit_total(node)
        return _pit_depth(node) / p if p != 0 else 0
    if n.startswith("cd_"):           # features.py:201-215
        i = int(n[3:])
        t = _tops(node)
        return t[i] - t[i - 1]
    if n == "all_diffs":
        t = _tops(node)
        return sum(t[i + 1] - t[i] for i in range(COLS - 1))
    if n == "max_diffs":
        t = _tops(node)
        return max(t[i + 1] - t[i] for i in range(COLS - 1))
    if n == "jaggedness":
        t = _tops(node)
        return sum(abs(t[i + 1] - t[i]) for i in range(COLS - 1))
    if n == "d_max_ht":               # features.py:239-257: same feature on state.prev, subtract
        return max(_tops(node)) - max(_tops(node.prev))
    if n == "d_all_ht":
        return sum(_tops(node)) - sum(_tops(node.prev))
    if n == "d_mean_ht":
        return _mean_ht(node) - _mean_ht(node.prev)
    if n == "d_pits":
        return _pit_total(node) - _pit_total(node.prev)
    if n == "landing_height":         # features.py:265-267 (pre-clear coordinates)
        return ROWS - (node.row + PIECES[node.piece][node.rot].shape[0])
    if n == "pattern_div":            # features.py:271-273
        d = np.diff(node.grid.cells.astype(int))
        return len({tuple(d[:, j]) for j in range(d.shape[1])})
    if n == "nine_filled":
        return len(_nine(node))
    if n == "tetris_progress":
        return _tetris_progress(node)
    if n == "full_cells":             # features.py:323-325
        return int(np.sum(node.grid.cells))
    if n == "weighted_cells":         # features.py:329-331
        return int(np.sum(np.sum(node.grid.cells, axis=1) * np.arange(ROWS, 0, -1)))
    if n == "eroded_cells":
        return _eroded(node)
    if n == "cuml_eroded":            # features.py:345-347
        e = _eroded(node)
        return e * (e + 1) / 2
    raise KeyError(fid)


def all_features(node):
    return [feature(node, i) for i in range(NUM_FEATURES)]


# ---------------------------------------------------------------------------------------------
# Canonical harness pieces that the reference leaves to the caller.
# ---------------------------------------------------------------------------------------------
def linear_score(node, fids, weights):
    """acc = 0.0; acc += float(f) * w in the caller's feature order — one rounded multiply and one
    rounded add per feature (no fused multiply-add, no compensated sum)."""
    acc = 0.0
    for fid, w in zip(fids, weights):
        acc += float(feature(node, fid)) * w
    return acc


def argmax_first(nodes, key):
    """simulator.py:123-150 (policy_best/all_max) with tie_breaker = first of the max group."""
    best = None
    best_k = None
    for e in nodes:
        k = key(e)
        if best_k is None or best_k < k:
            best, best_k = e, k
    return best


def piece_sequence(seed, T):
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


def hashed_piece(seed, game, t):
    """Counter-based piece stream (K4 twin): bit-identical to tb_hashed_piece() in the CUDA core."""
    M = 0xFFFFFFFFFFFFFFFF
    x = (seed * 0x9E3779B97F4A7C15 + game * 0xD1B54A32D192ED03 + t * 0x8CB92BA72F3D8DD7 + 0x2545F4914F6CDD1D) & M
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & M
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & M
    x ^= x >> 31
    return ((x >> 32) * 7) >> 32


def simulate(node, pieces, fids, weights, lookahead=1, trace=None, prune=True, move_set="drop", pressure=None):
    """simulator.py:152-189 with move_gen=move_drop and the canonical linear policy.

    `pieces` is an iterable of piece ids; returns the final node and the number of pieces placed.
    """
    assert lookahead > 0
    move_gen = MOVE_GENERATORS[move_set]
    if pressure is not None:
        move_gen = pressure_filter(move_gen, **pressure)
    it = iter(pieces)
    key = lambda s: linear_score(s, fids, weights)     # noqa: E731
    placed = 0
    while True:
        visible = tuple(itertools.islice(it, 0, lookahead))           # :156
        layers = [[node]]
        for p in visible:                                             # :159-167
            layers.append([prev.child(p, *mv) for prev in layers[-1] for mv in move_gen(prev, p)])
        nxt = None
        while nxt is None and len(layers) > 1:                        # :170-174
            pool = layers.pop()
            if pool:
                nxt = argmax_first(pool, key)
        if nxt is None:                                               # :187-189
            break
        while nxt.prev is not node:                                   # :178-179
            nxt = nxt.prev
        if trace is not None:
            trace += bytes([nxt.rot, nxt.row, nxt.col, len(nxt.gone)])
        placed += 1
        node = nxt
        if prune and node.prev is not None:
            node.prev.prev = None                                     # features look one level back only
        if len(visible) > 1:                                          # :185-186
            it = itertools.chain(visible[1:], it)
    return node, placed


def play_game(fids, weights, pieces, lookahead=1, want_trace=False, move_set="drop", pressure=None):
    """One canonical game from the empty board -> dict(pieces, lines, hist[4], score, trace)."""
    trace = bytearray() if want_trace else None
    root = Node(None, Grid())
    last, placed = simulate(root, pieces, fids, weights, lookahead, trace, prune=not want_trace, move_set=move_set, pressure=pressure)
    out = {"pieces": placed, "lines": last.lines(), "hist": [last.hist.get(k, 0) for k in (1, 2, 3, 4)]}
    if want_trace:
        out["score"] = game_score(last)
        out["trace"] = bytes(trace)
    return out


def candidate_fitness(fids, weights, sequences, lookahead=1):
    """fitness = total lines over the candidate's games / number of games."""
    total = 0
    for seq in sequences:
        res = play_game(fids, weights, seq, lookahead)
        total += res["lines"]
    return total / len(sequences)


# ---------------------------------------------------------------------------------------------
# Cross-entropy driver.  cogworks/learning.py:7-49
# ---------------------------------------------------------------------------------------------
def cross_entropy(feats, stdev, width, keep, rng, test_f, noise_f=None, map_f=map):
    if isinstance(feats, collections.abc.Mapping):                    # :9-16
        names, means = zip(*feats.items())
        means = list(means)
    else:
        names = list(feats)
        means = [0] * len(names)
    if isinstance(stdev, collections.abc.Mapping):                    # :19-24
        sig = [stdev[n] for n in names]
    else:
        sig = [stdev] * len(names)
    nf = len(names)
    while True:
        if noise_f is not None:                                       # :27-29
            sig = [noise_f(s) for s in sig]
        population = [[rng.gauss(means[i], sig[i]) for i in range(nf)] for _ in range(width)]   # :32-35
        results = map_f(lambda i: (test_f(dict(zip(names, population[i]))), i), range(width))  # :38
        elite = [population[i] for (_, i) in heapq.nlargest(keep, results)]                    # :41
        cols = list(zip(*elite))
        for i in range(nf):                                           # :44-46
            means[i] = np.mean(cols[i])
            sig[i] = np.std(cols[i])
        yield (collections.OrderedDict(zip(names, means)), collections.OrderedDict(zip(names, sig)))   # :49
