"""Engine — thin Python object over the C ABI (one ctx per device).

Host-buffer calls take/return numpy arrays (the library stages them; the call is synchronous).
Device-buffer calls take raw device pointers (e.g. torch tensors' data_ptr()) and a CUDA stream.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .tetris.features import FEATURE_ID, FEATURE_NAMES


class EngineError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data if a is not None else None


def _set_pressure(a, pressure):
    """pressure: None or a dict with any of avg_lat / resp_time / move_eff / two_but_rot (reference defaults, simulator.py:88)."""
    if pressure is None:
        return
    a.move_set |= _lib.MOVES_PRESSURE
    a.pressure_avg_lat = float(pressure.get("avg_lat", 250.9))
    a.pressure_resp_time = float(pressure.get("resp_time", 79.8))
    a.pressure_move_eff = float(pressure.get("move_eff", 1.23))
    a.pressure_two_but_rot = 1 if pressure.get("two_but_rot", False) else 0


def _set_points(a, points):
    """points: None (the reference's default table) or State.score's `points` argument (game.py:159): five non-negative ints."""
    if points is None:
        return
    pts = [int(x) for x in points]
    if len(pts) != 5 or any(x < 0 or x >= 2 ** 32 or x != y for x, y in zip(pts, points)):
        raise EngineError("points must be five non-negative integers below 2**32")
    if not any(pts):
        raise EngineError("an all-zero points table cannot be expressed at the C ABI (it selects the default table)")
    for i, x in enumerate(pts):
        a.points[i] = x


def feature_ids(features):
    """Accepts ids, names, or feature objects carrying __name__ (the reference's protocol, feature.py:23)."""
    out = []
    for f in features:
        if isinstance(f, (int, np.integer)):
            out.append(int(f))
        elif isinstance(f, str):
            out.append(FEATURE_ID[f])
        else:
            out.append(FEATURE_ID[f.__name__])
    return out


class Engine:
    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = ctypes.c_void_p()
        rc = self._L.tb200_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise EngineError("tb200_create(device=%d) failed with code %d: no usable CUDA device "
                              "(this engine has no CPU fallback)" % (device, rc))
        self._h = h
        self.device = int(device)
        self.features = None
        self._last_features = None
        sms, khz = ctypes.c_int32(), ctypes.c_int32()
        name = ctypes.create_string_buffer(128)
        self._check(self._L.tb200_device_info(self._h, ctypes.byref(sms), ctypes.byref(khz), name, 128))
        self.sms, self.clock_khz, self.device_name = sms.value, khz.value, name.value.decode()

    def close(self):
        """Destroy the ctx.  Pinned buffers handed out by pinned_empty() are NOT freed here: each one is released when the
        last numpy view of it is garbage-collected (weakref.finalize), so arrays that outlive the Engine stay valid."""
        if getattr(self, "_h", None):
            self._L.tb200_destroy(self._h)
            self._h = None

    def pinned_empty(self, shape, dtype):
        """numpy array over page-locked host memory.  The allocation lives exactly as long as the arrays that view it (the
        base object carries a finalizer), independent of this Engine.  Repeating a play_games call on the same pinned
        buffers lets the library replay it as a CUDA graph."""
        import weakref
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = ctypes.c_void_p()
        if self._L.tb200_alloc_pinned(max(n, 1), ctypes.byref(p)) != 0:
            raise EngineError("tb200_alloc_pinned(%d) failed" % n)
        buf = (ctypes.c_char * max(n, 1)).from_address(p.value)
        weakref.finalize(buf, self._L.tb200_free_pinned, ctypes.c_void_p(p.value))      # numpy keeps `buf` alive as the arrays' base
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EngineError("libtetris_b200 error %d: %s" % (rc, self._L.tb200_last_error(self._h).decode()))

    # ------------------------------------------------------------------ scorer definition
    def set_features(self, features):
        key = tuple(features)
        if key == self._last_features:                # same list as last time (evaluators call this every generation)
            return self
        ids = np.asarray(feature_ids(features), dtype=np.int32)
        self._check(self._L.tb200_set_features(self._h, _ptr(ids), len(ids)))
        self.features = [int(i) for i in ids]
        self._last_features = key
        return self

    @property
    def F(self):
        if self.features is None:
            raise EngineError("call set_features() first")
        return len(self.features)

    # ------------------------------------------------------------------ K2 whole games
    def play_games(self, weights, games_per_candidate, pieces=None, T=None, seq_of_game=None, seed=0, lookahead=1,
                   lanes=0, max_moves=0, init_rows=None, want_trace=False, want_scores=False, want_final=False,
                   want_fitness=True, out_results=None, out_fitness=None, move_set=0, pressure=None, init_lines=None, points=None):
        """weights f64[C][F] -> dict(results=structured[C*G], fitness=f64[C], trace=..., ...)   (host buffers)."""
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.ndim == 1:
            w = w[None, :]
        C, F = w.shape
        if F != self.F:
            raise EngineError("weights have %d columns, %d features set" % (F, self.F))
        G = int(games_per_candidate)
        n = C * G
        S = 0
        if pieces is not None:
            pieces = np.ascontiguousarray(pieces, dtype=np.uint8)
            if pieces.ndim == 1:
                pieces = pieces[None, :]
            S, T = pieces.shape
        elif T is None:
            raise EngineError("T is required with the counter-based piece stream")
        if seq_of_game is not None:
            seq_of_game = np.ascontiguousarray(seq_of_game, dtype=np.int32)
            if seq_of_game.size != n:
                raise EngineError("seq_of_game must have C*G entries")
        if init_rows is not None:
            init_rows = np.ascontiguousarray(init_rows, dtype=np.uint16).reshape(n, 20)
        T = int(T)
        results = out_results if out_results is not None else np.zeros(n, dtype=_lib.RESULT_DTYPE)
        if results.shape != (n,) or results.dtype != _lib.RESULT_DTYPE:
            raise EngineError("out_results must be a structured array of C*G records")
        trace = np.zeros((n, T, 4), dtype=np.uint8) if want_trace else None
        scores = np.zeros((n, T), dtype=np.float64) if want_scores else None
        final = np.zeros((n, 20), dtype=np.uint16) if want_final else None
        fitness = (out_fitness if out_fitness is not None else np.zeros(C, dtype=np.float64)) if want_fitness else None
        a = _lib.PlayArgs()
        a.struct_size = ctypes.sizeof(_lib.PlayArgs)
        a.mem = _lib.MEM_HOST
        a.stream = None
        a.n_candidates, a.games_per_candidate = C, G
        a.weights, a.pieces, a.n_sequences = _ptr(w), _ptr(pieces), S
        a.T, a.max_moves, a.seq_of_game, a.seed = T, int(max_moves), _ptr(seq_of_game), int(seed)
        a.lookahead, a.lanes, a.init_rows = int(lookahead), int(lanes), _ptr(init_rows)
        a.results, a.trace, a.score_trace = _ptr(results), _ptr(trace), _ptr(scores)
        a.final_rows, a.fitness = _ptr(final), _ptr(fitness)
        a.move_set = int(move_set)
        _set_pressure(a, pressure)
        _set_points(a, points)
        if init_lines is not None:
            init_lines = np.ascontiguousarray(np.broadcast_to(np.asarray(init_lines, dtype=np.uint32), (n,)), dtype=np.uint32)
            a.init_lines = _ptr(init_lines)
        self._check(self._L.tb200_play_games(self._h, ctypes.byref(a)))
        return {"results": results, "fitness": fitness, "trace": trace, "scores": scores, "final": final}

    def play_games_device(self, stream, n_candidates, games_per_candidate, weights_ptr, results_ptr, T, pieces_ptr=0,
                          n_sequences=0, seq_of_game_ptr=0, seed=0, lookahead=1, lanes=0, max_moves=0, init_rows_ptr=0,
                          trace_ptr=0, score_trace_ptr=0, final_rows_ptr=0, fitness_ptr=0, move_set=0, points=None):
        """All pointers are DEVICE pointers (ints); work is enqueued on `stream` (a cudaStream_t int) and not synchronised.
        One call in flight per Engine: the library orders a later call behind this one (see include/tetris_b200.h);
        check_status() waits for it and reports an expired device-side wait."""
        self.run_prepared(self.prepare_device_call(stream, n_candidates, games_per_candidate, weights_ptr, results_ptr, T, pieces_ptr, n_sequences,
                                                   seq_of_game_ptr, seed, lookahead, lanes, max_moves, init_rows_ptr, trace_ptr, score_trace_ptr,
                                                   final_rows_ptr, fitness_ptr, move_set, points))

    def run_prepared(self, args):
        """Enqueue a device-mode call whose argument block was built once by prepare_device_call (latency-sensitive loops)."""
        self._check(self._L.tb200_play_games(self._h, ctypes.byref(args)))

    def prepare_device_call(self, stream, n_candidates, games_per_candidate, weights_ptr, results_ptr, T, pieces_ptr=0,
                            n_sequences=0, seq_of_game_ptr=0, seed=0, lookahead=1, lanes=0, max_moves=0, init_rows_ptr=0,
                            trace_ptr=0, score_trace_ptr=0, final_rows_ptr=0, fitness_ptr=0, move_set=0, points=None):
        a = _lib.PlayArgs()
        a.struct_size = ctypes.sizeof(_lib.PlayArgs)
        a.mem = _lib.MEM_DEVICE
        a.stream = stream or None
        a.n_candidates, a.games_per_candidate = int(n_candidates), int(games_per_candidate)
        a.weights, a.pieces, a.n_sequences = weights_ptr or None, pieces_ptr or None, int(n_sequences)
        a.T, a.max_moves, a.seq_of_game, a.seed = int(T), int(max_moves), seq_of_game_ptr or None, int(seed)
        a.lookahead, a.lanes, a.init_rows = int(lookahead), int(lanes), init_rows_ptr or None
        a.results, a.trace, a.score_trace = results_ptr or None, trace_ptr or None, score_trace_ptr or None
        a.final_rows, a.fitness = final_rows_ptr or None, fitness_ptr or None
        a.move_set = int(move_set)
        _set_points(a, points)
        return a

    def check_status(self):
        """Wait for the last device-mode call and raise if one of its bounded device-side waits expired."""
        self._check(self._L.tb200_check_status(self._h))

    def last_timing(self):
        t = _lib.Timing()
        self._check(self._L.tb200_last_timing(self._h, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in _lib.Timing._fields_}

    # ------------------------------------------------------------------ K1 one decision per board
    def best_move(self, boards, piece, next_piece=None, weights=None, lookahead=1, want_rows=False, move_set=0, pressure=None, lines=None):
        """One placement decision per board.  move_set / pressure as in play_games; lines = lines cleared so far on each
        board (the level the pressure filter reads)."""
        boards = np.ascontiguousarray(boards, dtype=np.uint16).reshape(-1, 20)
        N = boards.shape[0]
        piece = np.ascontiguousarray(piece, dtype=np.uint8).reshape(N)
        nxt = np.ascontiguousarray(next_piece, dtype=np.uint8).reshape(N) if next_piece is not None else None
        w = np.ascontiguousarray(weights, dtype=np.float64)
        per_board = 1 if (w.ndim == 2 and w.shape[0] == N and N > 1) else 0
        if w.reshape(-1, self.F).shape[0] not in (1, N):
            raise EngineError("weights must be f64[F] or f64[N][F]")
        move = np.zeros((N, 4), dtype=np.uint8)
        score = np.zeros(N, dtype=np.float64)
        count = np.zeros(N, dtype=np.uint32)
        rows = np.zeros((N, 20), dtype=np.uint16) if want_rows else None
        opts = _lib.MoveOpts()
        opts.struct_size = ctypes.sizeof(_lib.MoveOpts)
        opts.move_set = int(move_set)
        _set_pressure(opts, pressure)
        ln = np.ascontiguousarray(lines, dtype=np.uint32).reshape(N) if lines is not None else None
        opts.lines = _ptr(ln)
        self._check(self._L.tb200_best_move_ex(self._h, _lib.MEM_HOST, None, N, _ptr(boards), _ptr(piece), _ptr(nxt), _ptr(w),
                                               per_board, int(lookahead), ctypes.byref(opts), _ptr(move), _ptr(score), _ptr(count), _ptr(rows)))
        return {"move": move, "score": score, "count": count, "rows": rows}

    # ------------------------------------------------------------------ K3 features of placements
    def eval_features(self, prev_boards, piece, rot, col, want_rows=False, row=None):
        prev = np.ascontiguousarray(prev_boards, dtype=np.uint16).reshape(-1, 20)
        N = prev.shape[0]
        piece = np.ascontiguousarray(piece, dtype=np.uint8).reshape(N)
        rot = np.ascontiguousarray(rot, dtype=np.uint8).reshape(N)
        col = np.ascontiguousarray(col, dtype=np.uint8).reshape(N)
        at_row = np.ascontiguousarray(row, dtype=np.uint8).reshape(N) if row is not None else None
        feats = np.zeros((N, _lib.NUM_FEATURES), dtype=np.float64)
        info = np.zeros((N, 3), dtype=np.int32)
        rows = np.zeros((N, 20), dtype=np.uint16) if want_rows else None
        self._check(self._L.tb200_eval_features_at(self._h, _lib.MEM_HOST, None, N, _ptr(prev), _ptr(piece), _ptr(rot), _ptr(col),
                                                   _ptr(at_row), _ptr(feats), _ptr(info), _ptr(rows)))
        return {"features": feats, "legal": info[:, 0], "row": info[:, 1], "k": info[:, 2], "rows": rows}

    # ------------------------------------------------------------------ K4 piece stream
    def hashed_pieces(self, seed, first_stream, n_streams, T):
        out = np.zeros((int(n_streams), int(T)), dtype=np.uint8)
        self._check(self._L.tb200_hashed_pieces(self._h, _lib.MEM_HOST, None, int(seed), int(first_stream), int(n_streams), int(T), _ptr(out)))
        return out


__all__ = ["Engine", "EngineError", "feature_ids", "FEATURE_NAMES", "FEATURE_ID"]
