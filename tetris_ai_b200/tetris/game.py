"""Host-side game model with the reference's names (cogworks/tetris/game.py): Zoid, zoids.classic,
Board, State (+ State.Delta).  It exists so that code written against the reference keeps working
around the GPU path: boards cross the C ABI as 20 row masks, and the states that the GPU
simulate() yields are rebuilt from the device's move trace with Board.imprint/clear (a replay of
already-chosen moves, not a search).

Unlike the reference (a numpy bool grid), a Board here stores one 10-bit mask per row, top row
first — the exchange format of include/tetris_b200.h; `Board.data` materialises the bool grid.
"""
from __future__ import annotations

import collections

import numpy as np

_SHAPES = collections.OrderedDict([          # base orientation, rows top to bottom (game.py:24-34)
    ("I", ("####",)), ("T", ("###", ".#.")), ("L", ("###", "#..")), ("J", ("###", "..#")),
    ("O", ("##", "##")), ("Z", ("##.", ".##")), ("S", (".##", "##.")),
])
_ROTATIONS = {"I": 2, "T": 4, "L": 4, "J": 4, "O": 1, "Z": 2, "S": 2}


def _quarter_turn(rows):
    h, w = len(rows), len(rows[0])
    return tuple("".join(rows[h - 1 - j][i] for j in range(h)) for i in range(w))


class Zoid:
    """A piece: zoid[rot] is the bool array of rotation `rot` (clockwise quarter turns, game.py:10)."""

    def __init__(self, name, rows, rots, pid):
        self.name = name
        self.id = pid
        shapes, masks = [], []
        for _ in range(rots):
            shapes.append(np.array([[ch == "#" for ch in r] for r in rows], dtype=np.bool_))
            masks.append(tuple(sum(1 << j for j, ch in enumerate(r) if ch == "#") for r in rows))
            rows = _quarter_turn(rows)
        self.shapes = tuple(shapes)
        self.row_masks = tuple(masks)            # per rotation: one bit mask per piece row (bit j = piece column j)

    def __repr__(self):
        return "Zoid.%s" % self.name

    def __len__(self):
        return len(self.shapes)

    def __getitem__(self, rot):
        return self.shapes[rot]


_Classic = collections.namedtuple("classic", list(_SHAPES))
_Zoids = collections.namedtuple("zoids", ["classic"])
zoids = _Zoids(classic=_Classic(*[Zoid(n, rows, _ROTATIONS[n], i) for i, (n, rows) in enumerate(_SHAPES.items())]))
PIECES = tuple(zoids.classic)                     # piece id -> Zoid (I,T,L,J,O,Z,S = 0..6)


class Board:
    def __init__(self, rows=20, cols=10, zero=True, masks=None):
        if rows < 1 or cols < 1:
            raise ValueError("a board needs at least one row and one column")
        self._nrows, self._ncols = rows, cols
        self.masks = list(masks) if masks is not None else [0] * rows
        self.heights = self._heights()

    # -- geometry -------------------------------------------------------------------------
    def rows(self):
        return self._nrows

    def cols(self):
        return self._ncols

    def _heights(self):
        out = [0] * self._ncols
        for c in range(self._ncols):
            bit = 1 << c
            for r, m in enumerate(self.masks):
                if m & bit:
                    out[c] = self._nrows - r
                    break
        return out

    def height(self, col):
        return self.heights[col]

    @property
    def data(self):
        return np.array([[(m >> c) & 1 for c in range(self._ncols)] for m in self.masks], dtype=np.bool_)

    def rows16(self):
        """The board in the exchange format of include/tetris_b200.h.  The host model (replay, scoring, pickling) works for
        any size like the reference's (game.py:38-45); the device kernels are specialised to the 20 x 10 board every
        reference config uses, so this — the only way a board reaches the device — refuses other sizes."""
        if (self._nrows, self._ncols) != (20, 10):
            raise ValueError("the B200 engine is specialised to the 20 x 10 board of the reference configs (this board is %d x %d)"
                             % (self._nrows, self._ncols))
        return np.array(self.masks, dtype=np.uint16)

    def row(self, r):
        return self.data[r]

    def col(self, c):
        return self.data[:, c]

    def __getitem__(self, pos):
        return self.data[pos]

    # -- mutation (replay of a chosen move) -----------------------------------------------------
    def overlaps(self, orient, row, col):
        h, w = orient.shape
        if row < 0 or col < 0 or row + h > self._nrows or col + w > self._ncols:
            return True
        for i in range(h):
            m = sum(1 << (col + j) for j in range(w) if orient[i, j])
            if self.masks[row + i] & m:
                return True
        return False

    def imprint(self, orient, row, col):
        h, w = orient.shape
        for i in range(h):
            self.masks[row + i] |= sum(1 << (col + j) for j in range(w) if orient[i, j])
        self.heights = self._heights()

    def full(self):
        everything = (1 << self._ncols) - 1
        return {r for r, m in enumerate(self.masks) if m == everything}

    def clear(self, rows):
        kept = [m for r, m in enumerate(self.masks) if r not in rows]
        self.masks = [0] * (self._nrows - len(kept)) + kept
        self.heights = self._heights()

    def copy(self):
        return Board(self._nrows, self._ncols, masks=self.masks)

    __deepcopy__ = lambda self, memo: self.copy()      # noqa: E731

    def __eq__(self, other):
        return self.masks == other.masks

    def __ne__(self, other):
        return self.masks != other.masks

    def __repr__(self):
        return "%dx%d Board@%#x" % (self._nrows, self._ncols, id(self))

    def __str__(self):
        return "\n".join("".join("#" if (m >> c) & 1 else "." for c in range(self._ncols)) for m in self.masks)


class State:
    """Immutable game node: (prev, board, cleared histogram, delta) as in game.py:110-127."""

    class Delta:
        __slots__ = ("zoid", "rot", "row", "col", "cleared")

        def __init__(self, zoid, rot, row, col, cleared):
            self.zoid, self.rot, self.row, self.col, self.cleared = zoid, rot, row, col, cleared

    def __init__(self, prev, board, cleared=None, delta=None):
        self.prev = prev
        self.board = board
        self.cleared = cleared if cleared is not None else collections.defaultdict(int)
        self.delta = delta

    def future(self, zoid, rot, row, col):
        """Apply one (already chosen) placement: copy, imprint, find full rows, clear (game.py:166-180)."""
        board = self.board.copy()
        hist = self.cleared.copy()
        board.imprint(zoid[rot], row, col)
        gone = frozenset(board.full())
        if gone:
            hist[len(gone)] += 1
            board.clear(gone)
        return State(self, board, cleared=hist, delta=State.Delta(zoid, rot, row, col, gone))

    def lines_cleared(self, simultaneously=None):
        if simultaneously is not None:
            return self.cleared[simultaneously]
        return sum(k * n for k, n in self.cleared.items())

    def level(self):
        return self.lines_cleared() // 10

    def score(self, points=(0, 40, 100, 300, 1200)):
        total, node = 0, self
        while node.delta is not None:
            total += points[len(node.delta.cleared)] * (node.level() + 1)
            node = node.prev
        return total

    # pickling = (root board, root histogram, move list); replayed on load (game.py:129-148)
    def __getstate__(self):
        moves, node = [], self
        while node.prev is not None:
            d = node.delta
            moves.append((d.zoid.id, d.rot, d.row, d.col))
            node = node.prev
        return (list(node.board.masks), dict(node.cleared), moves[::-1], (node.board.rows(), node.board.cols()))

    def __setstate__(self, saved):
        masks, hist, moves = saved[:3]
        rows, cols = saved[3] if len(saved) > 3 else (20, 10)
        node = State(None, Board(rows, cols, masks=masks), cleared=collections.defaultdict(int, hist))
        for pid, rot, row, col in moves:
            node = node.future(PIECES[pid], rot, row, col)
        self.prev, self.board, self.cleared, self.delta = node.prev, node.board, node.cleared, node.delta
