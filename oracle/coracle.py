"""ctypes wrapper around oracle/libcoracle.so (the plain-C restatement, coracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libcoracle.so")
    src = os.path.join(_HERE, "coracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libcoracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u16p = ctypes.POINTER(ctypes.c_uint16)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f64p = ctypes.POINTER(ctypes.c_double)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        L.co_drop_moves.argtypes = [u16p, ctypes.c_int, i32p]
        L.co_drop_moves.restype = ctypes.c_int
        L.co_eval_placement.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, f64p, u16p, i32p, i32p]
        L.co_eval_placement.restype = ctypes.c_int
        L.co_moves.argtypes = [u16p, ctypes.c_int, ctypes.c_int, i32p]
        L.co_moves.restype = ctypes.c_int
        L.co_best_move.argtypes = [u16p, ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, f64p, ctypes.c_int, i32p, f64p, ctypes.c_int]
        L.co_best_move.restype = ctypes.c_int
        L.co_hashed_piece.argtypes = [ctypes.c_uint64] * 3
        L.co_hashed_piece.restype = ctypes.c_int
        L.co_play_games.argtypes = [i32p, ctypes.c_int, f64p, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int,
                                    ctypes.c_uint32, i32p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, u64p, u8p, ctypes.c_int, f64p]
        L.co_play_games.restype = ctypes.c_int
        L.co_play_one.argtypes = [i32p, ctypes.c_int, f64p, u8p, ctypes.c_uint32, ctypes.c_int, u16p, u64p, u8p, u16p, ctypes.c_int, f64p]
        L.co_play_one.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def drop_moves(rows, piece):
    rows = np.ascontiguousarray(rows, dtype=np.uint16)
    mv = np.zeros((34, 3), dtype=np.int32)
    n = lib().co_drop_moves(_p(rows, ctypes.c_uint16), int(piece), _p(mv, ctypes.c_int32))
    return [tuple(int(x) for x in mv[i]) for i in range(n)]


MOVE_SETS = {"drop": 0, "slide": 1, "wiggle_top": 4, "wiggle_drop": 8, 0: 0, 1: 1, 4: 4, 8: 8}


def _pressure(pressure):
    """pressure: None or dict(avg_lat, resp_time, move_eff, two_but_rot) with the reference's defaults (simulator.py:88)."""
    if pressure is None:
        return 0, None
    a = np.array([pressure.get("avg_lat", 250.9), pressure.get("resp_time", 79.8), pressure.get("move_eff", 1.23),
                  1.0 if pressure.get("two_but_rot", False) else 0.0], dtype=np.float64)
    return 2, a


def moves(rows, piece, move_set="drop"):
    """move list of move_drop ('drop') or move_with_overhang_slide(move_drop) ('slide')."""
    rows = np.ascontiguousarray(rows, dtype=np.uint16)
    mv = np.zeros((640, 3), dtype=np.int32)
    n = lib().co_moves(_p(rows, ctypes.c_uint16), int(piece), MOVE_SETS[move_set], _p(mv, ctypes.c_int32))
    return [tuple(int(x) for x in mv[i]) for i in range(n)]


def eval_placement(prev_rows, piece, rot, col):
    """-> None if illegal else dict(features f64[45], post u16[20], row, k)."""
    prev_rows = np.ascontiguousarray(prev_rows, dtype=np.uint16)
    f = np.zeros(45, dtype=np.float64)
    post = np.zeros(20, dtype=np.uint16)
    row = ctypes.c_int32(0)
    k = ctypes.c_int32(0)
    ok = lib().co_eval_placement(_p(prev_rows, ctypes.c_uint16), int(piece), int(rot), int(col),
                                 _p(f, ctypes.c_double), _p(post, ctypes.c_uint16), ctypes.byref(row), ctypes.byref(k))
    if not ok:
        return None
    return {"features": f, "post": post, "row": row.value, "k": k.value}


def best_move(rows, piece, next_piece, lookahead, fids, weights, move_set="drop"):
    """-> None (no move) or dict(rot,row,col,k,score,n_scored)."""
    rows = np.ascontiguousarray(rows, dtype=np.uint16)
    fids = np.ascontiguousarray(fids, dtype=np.int32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    mv = np.zeros(4, dtype=np.int32)
    sc = ctypes.c_double(0.0)
    n = lib().co_best_move(_p(rows, ctypes.c_uint16), int(piece), int(next_piece), int(lookahead),
                           _p(fids, ctypes.c_int32), _p(w, ctypes.c_double), len(fids), _p(mv, ctypes.c_int32), ctypes.byref(sc),
                           MOVE_SETS[move_set])
    if n == 0:
        return None
    return {"rot": int(mv[0]), "row": int(mv[1]), "col": int(mv[2]), "k": int(mv[3]), "score": sc.value, "n_scored": n}


def hashed_piece(seed, game, t):
    return lib().co_hashed_piece(seed, game, t)


def play_games(fids, weights, G, pieces=None, T=None, seq_of_game=None, seed=0, lookahead=1, threads=1, want_trace=False,
               move_set="drop", pressure=None):
    """weights f64[C][F]; pieces u8[S][T] or None (hashed stream).  -> dict of u64 arrays [C*G] (+ trace u8[C*G][T][4])."""
    fids = np.ascontiguousarray(fids, dtype=np.int32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if w.ndim == 1:
        w = w[None, :]
    C, F = w.shape
    assert F == len(fids)
    if pieces is not None:
        pieces = np.ascontiguousarray(pieces, dtype=np.uint8)
        if pieces.ndim == 1:
            pieces = pieces[None, :]
        S, T = pieces.shape
    else:
        S = 0
        assert T is not None
    if seq_of_game is not None:
        seq_of_game = np.ascontiguousarray(seq_of_game, dtype=np.int32)
        assert seq_of_game.size == C * G
    res = np.zeros((C * G, 8), dtype=np.uint64)
    trace = np.zeros((C * G, T, 4), dtype=np.uint8) if want_trace else None
    rc = lib().co_play_games(_p(fids, ctypes.c_int32), F, _p(w, ctypes.c_double), C, G, _p(pieces, ctypes.c_uint8), S,
                             int(T), _p(seq_of_game, ctypes.c_int32), int(seed), int(lookahead), int(threads),
                             _p(res, ctypes.c_uint64), _p(trace, ctypes.c_uint8), MOVE_SETS[move_set] | _pressure(pressure)[0],
                             _p(_pressure(pressure)[1], ctypes.c_double))
    assert rc == 0
    out = {"pieces": res[:, 0], "lines": res[:, 1], "hist": res[:, 2:6], "score": res[:, 6], "placements": res[:, 7]}
    if want_trace:
        out["trace"] = trace
    return out


def play_one(fids, weights, pieces, lookahead=1, init_rows=None, move_set="drop", pressure=None, init_lines=0):
    fids = np.ascontiguousarray(fids, dtype=np.int32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    pieces = np.ascontiguousarray(pieces, dtype=np.uint8)
    T = len(pieces)
    init = np.ascontiguousarray(init_rows, dtype=np.uint16) if init_rows is not None else None
    res = np.zeros(8, dtype=np.uint64)
    trace = np.zeros((max(T, 1), 4), dtype=np.uint8)
    final = np.zeros(20, dtype=np.uint16)
    lib().co_play_one_ex(_p(fids, ctypes.c_int32), len(fids), _p(w, ctypes.c_double), _p(pieces, ctypes.c_uint8), T, int(lookahead),
                         _p(init, ctypes.c_uint16), ctypes.c_uint32(int(init_lines)), _p(res, ctypes.c_uint64), _p(trace, ctypes.c_uint8),
                         _p(final, ctypes.c_uint16), MOVE_SETS[move_set] | _pressure(pressure)[0], _p(_pressure(pressure)[1], ctypes.c_double))
    n = int(res[0])
    return {"pieces": n, "lines": int(res[1]), "hist": [int(x) for x in res[2:6]], "score": int(res[6]),
            "placements": int(res[7]), "trace": trace[:n].copy(), "final": final}
