"""ctypes access to tests/emu/libemu.so — the HOST build of the CUDA core (test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "emu", "emu.cpp")
_SO = os.path.join(_HERE, "emu", "libemu.so")
_CORE = os.path.join(_HERE, "..", "tetris_ai_b200", "csrc", "tetris_core.cuh")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        newest = max(os.path.getmtime(_SRC), os.path.getmtime(_CORE))
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-o", _SO, _SRC])
        L = ctypes.CDLL(_SO)
        L.emu_play_one.restype = ctypes.c_int
        L.emu_hashed_piece.restype = ctypes.c_int
        L.emu_hashed_piece.argtypes = [ctypes.c_uint64] * 3
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def play_one(fids, weights, pieces=None, T=None, seed=0, stream=0, lookahead=1, max_moves=0, mode=-1, init_rows=None, move_set=0, pressure=None):
    fids = np.ascontiguousarray(fids, dtype=np.int32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if pieces is not None:
        pieces = np.ascontiguousarray(pieces, dtype=np.uint8)
        T = len(pieces)
    init = np.ascontiguousarray(init_rows, dtype=np.uint16) if init_rows is not None else None
    res = np.zeros(8, dtype=np.uint64)
    trace = np.zeros((max(T, 1), 4), dtype=np.uint8)
    st = np.zeros(max(T, 1), dtype=np.float64)
    final = np.zeros(20, dtype=np.uint16)
    press = None
    if pressure is not None:
        press = np.array([pressure.get("avg_lat", 250.9), pressure.get("resp_time", 79.8), pressure.get("move_eff", 1.23),
                          1.0 if pressure.get("two_but_rot", False) else 0.0], dtype=np.float64)
        move_set |= 2
    m = lib().emu_play_one(_p(fids), ctypes.c_int(len(fids)), _p(w), _p(pieces), ctypes.c_uint32(T), ctypes.c_uint64(seed),
                           ctypes.c_uint64(stream), ctypes.c_int(lookahead), ctypes.c_uint32(max_moves), ctypes.c_int(mode),
                           _p(init), _p(res), _p(trace), _p(st), _p(final), ctypes.c_int(move_set), _p(press))
    n = int(res[0])
    return {"pieces": n, "lines": int(res[1]), "hist": [int(x) for x in res[2:6]], "score": int(res[6]),
            "placements": int(res[7]), "trace": trace[:n].copy(), "scores": st[:n].copy(), "final": final, "mode": m}


def eval_placement(prev_rows, piece, rot, col):
    prev = np.ascontiguousarray(prev_rows, dtype=np.uint16)
    f = np.zeros(45, dtype=np.float64)
    post = np.zeros(20, dtype=np.uint16)
    info = np.zeros(3, dtype=np.int32)
    lib().emu_eval_placement(_p(prev), ctypes.c_int(piece), ctypes.c_int(rot), ctypes.c_int(col), _p(f), _p(post), _p(info))
    return {"legal": int(info[0]), "row": int(info[1]), "k": int(info[2]), "features": f, "post": post}


def moves(rows, piece, move_set=0):
    rows = np.ascontiguousarray(rows, dtype=np.uint16)
    out = np.zeros((640, 3), dtype=np.int32)
    lib().emu_moves.restype = ctypes.c_int
    n = lib().emu_moves(_p(rows), ctypes.c_int(piece), ctypes.c_int(move_set), _p(out))
    return [tuple(int(x) for x in out[i]) for i in range(n)]


def wiggle_set(rows, piece, from_top=True):
    rows = np.ascontiguousarray(rows, dtype=np.uint16)
    out = np.zeros((640, 3), dtype=np.int32)
    lib().emu_wiggle_set.restype = ctypes.c_int
    n = lib().emu_wiggle_set(_p(rows), ctypes.c_int(piece), ctypes.c_int(1 if from_top else 0), _p(out))
    return sorted(tuple(int(x) for x in out[i]) for i in range(n))


def wiggle_stop_check(rows, piece, from_top, seed, trials=8):
    rows = np.ascontiguousarray(rows, dtype=np.uint16)
    lib().emu_wiggle_stop_check.restype = ctypes.c_int
    return lib().emu_wiggle_stop_check(_p(rows), ctypes.c_int(piece), ctypes.c_int(1 if from_top else 0), ctypes.c_uint32(seed), ctypes.c_int(trials))
