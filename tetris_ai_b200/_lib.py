"""ctypes binding of libtetris_b200.so (the C ABI declared in include/tetris_b200.h).

There is no CPU fallback: if the shared library has not been built, or no CUDA device is usable,
the loader / Engine constructor raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TB200_LIB") or os.path.join(_HERE, "libtetris_b200.so")      # TB200_LIB: an experiment build of the same library

MEM_HOST, MEM_DEVICE = 0, 1
MOVES_DROP, MOVES_OVERHANG_SLIDE, MOVES_PRESSURE, MOVES_WIGGLE, MOVES_WIGGLE_DROP = 0, 1, 2, 4, 8
NUM_FEATURES = 45
MAX_WEIGHTED = 48

RESULT_DTYPE = np.dtype([("pieces", "<u4"), ("lines", "<u4"), ("hist", "<u4", (4,)), ("score", "<u8"), ("placements", "<u8")])
assert RESULT_DTYPE.itemsize == 40


class PlayArgs(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("mem", ctypes.c_int32), ("stream", ctypes.c_void_p),
        ("n_candidates", ctypes.c_int32), ("games_per_candidate", ctypes.c_int32),
        ("weights", ctypes.c_void_p), ("pieces", ctypes.c_void_p), ("n_sequences", ctypes.c_int32),
        ("T", ctypes.c_uint32), ("max_moves", ctypes.c_uint32), ("seq_of_game", ctypes.c_void_p),
        ("seed", ctypes.c_uint64), ("lookahead", ctypes.c_int32), ("lanes", ctypes.c_int32),
        ("init_rows", ctypes.c_void_p), ("results", ctypes.c_void_p), ("trace", ctypes.c_void_p),
        ("score_trace", ctypes.c_void_p), ("final_rows", ctypes.c_void_p), ("fitness", ctypes.c_void_p),
        ("move_set", ctypes.c_int32), ("pressure_two_but_rot", ctypes.c_int32),
        ("pressure_avg_lat", ctypes.c_double), ("pressure_resp_time", ctypes.c_double), ("pressure_move_eff", ctypes.c_double),
        ("init_lines", ctypes.c_void_p),
        ("points", ctypes.c_uint32 * 5), ("reserved0", ctypes.c_uint32),
    ]


class MoveOpts(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("move_set", ctypes.c_int32), ("pressure_two_but_rot", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("pressure_avg_lat", ctypes.c_double), ("pressure_resp_time", ctypes.c_double), ("pressure_move_eff", ctypes.c_double),
        ("lines", ctypes.c_void_p),
    ]


class Timing(ctypes.Structure):
    _fields_ = [
        ("kernel_ms", ctypes.c_float), ("total_ms", ctypes.c_float), ("lanes", ctypes.c_int32),
        ("grid", ctypes.c_int32), ("block", ctypes.c_int32), ("kernels_launched", ctypes.c_int32),
        ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64),
        ("specialised", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


EXPORTS = (
    "tb200_create", "tb200_destroy", "tb200_last_error", "tb200_version", "tb200_device_info",
    "tb200_set_features", "tb200_play_games", "tb200_last_timing", "tb200_best_move", "tb200_best_move_ex",
    "tb200_eval_features", "tb200_eval_features_at", "tb200_hashed_pieces", "tb200_alloc_pinned", "tb200_free_pinned",
    "tb200_check_status", "tb200_jit_build",
)

_LIB = None


def load():
    """Load the CUDA library; raise (never fall back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "tetris_ai_b200: %s has not been built (run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C tetris_ai_b200/csrc`); there is no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u32, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint64
    L.tb200_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.tb200_create.restype = ctypes.c_int
    L.tb200_destroy.argtypes = [vp]
    L.tb200_destroy.restype = None
    L.tb200_last_error.argtypes = [vp]
    L.tb200_last_error.restype = ctypes.c_char_p
    L.tb200_version.argtypes = []
    L.tb200_version.restype = ctypes.c_char_p
    L.tb200_device_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.c_char_p, i32]
    L.tb200_device_info.restype = ctypes.c_int
    L.tb200_set_features.argtypes = [vp, vp, i32]
    L.tb200_set_features.restype = ctypes.c_int
    L.tb200_play_games.argtypes = [vp, ctypes.POINTER(PlayArgs)]
    L.tb200_play_games.restype = ctypes.c_int
    L.tb200_last_timing.argtypes = [vp, ctypes.POINTER(Timing)]
    L.tb200_last_timing.restype = ctypes.c_int
    L.tb200_jit_build.argtypes = [vp, i32, ctypes.c_char_p, i32]
    L.tb200_jit_build.restype = ctypes.c_int
    L.tb200_check_status.argtypes = [vp]
    L.tb200_check_status.restype = ctypes.c_int
    L.tb200_best_move.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.tb200_best_move.restype = ctypes.c_int
    L.tb200_best_move_ex.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, i32, i32, ctypes.POINTER(MoveOpts), vp, vp, vp, vp]
    L.tb200_best_move_ex.restype = ctypes.c_int
    L.tb200_eval_features.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.tb200_eval_features.restype = ctypes.c_int
    L.tb200_eval_features_at.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.tb200_eval_features_at.restype = ctypes.c_int
    L.tb200_hashed_pieces.argtypes = [vp, i32, vp, u64, u64, i32, u32, vp]
    L.tb200_hashed_pieces.restype = ctypes.c_int
    L.tb200_alloc_pinned.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    L.tb200_alloc_pinned.restype = ctypes.c_int
    L.tb200_free_pinned.argtypes = [vp]
    L.tb200_free_pinned.restype = None
    _LIB = L
    return L
