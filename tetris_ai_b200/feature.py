"""Host-side mirror of cogworks/feature.py (the feature mini-framework) for the GPU path.

Same call signatures as the reference:
  define(*needs)                      feature.py:14   memoising decorator for host-side (Python) features
  with_transformed_state(fn, feature) feature.py:43
  evaluate(state, features)           feature.py:50   {feature: value} or {feature: value * weight}

The 45 Tetris features (tetris_ai_b200.tetris.features) are device handles: evaluate() sends the
state's (previous board, placement) to the K3 kernel and returns all requested values in one
launch.  Python features built with define() are still supported by evaluate() (they run as plain
Python, exactly like the reference), but they cannot be part of a GPU scoring policy.
"""
from __future__ import annotations

import collections.abc
import functools

from .tetris.features import FLOAT_FEATURES, Feature

_ENGINE = None


def default_engine():
    """Process-wide Engine on device 0, created on first use (raises when there is no GPU)."""
    global _ENGINE
    if _ENGINE is None:
        from .engine import Engine
        _ENGINE = Engine(0)
    return _ENGINE


def set_default_engine(engine):
    global _ENGINE
    _ENGINE = engine


def define(*needs):
    """Decorator factory: impl(state, *values_of_needs) -> feature(state, cache=None) with a per-evaluation
    memo keyed by the implementation and hit/miss counters (feature.accesses / feature.misses)."""
    def decorate(impl, name=None):
        if not callable(impl):
            raise TypeError("feature implementation must be callable")
        if name is not None:
            impl.__name__ = name

        @functools.wraps(impl)
        def wrapper(state, cache=None):
            memo = {} if cache is None else cache
            wrapper.accesses += 1
            if impl not in memo:
                wrapper.misses += 1
                memo[impl] = impl(state, *[dep(state, cache=memo) for dep in needs])
            return memo[impl]
        wrapper.accesses = 0
        wrapper.misses = 0
        return wrapper
    return decorate


def with_transformed_state(transformer, feature):
    """Evaluate `feature` on transformer(state); the cache is deliberately not forwarded (feature.py:44)."""
    def moved(state, cache=None):
        return feature(transformer(state))
    return moved


def _typed(name, value, pits):
    if name == "mean_pit_depth" and pits == 0:
        return 0                                   # features.py:194 returns the int 0
    return float(value) if name in FLOAT_FEATURES else int(value)


def device_features(state, engine=None):
    """All 45 device features of `state` (which must have been produced by a placement): one K3 launch."""
    if state.delta is None or state.prev is None:
        raise ValueError("device features need a state with a previous state and a delta (a placed piece)")
    eng = engine or default_engine()
    d = state.delta
    r = eng.eval_features([state.prev.board.rows16()], [d.zoid.id], [d.rot], [d.col], row=[d.row])
    if int(r["legal"][0]) != 1:
        raise ValueError("state.delta does not fit on its previous state's board")
    return r["features"][0]


_DEVICE_VALUES = "tetris_ai_b200.device_values"       # cache key under which one K3 launch's 45 values are memoised


def feature_value(f, state, cache=None, engine=None):
    """One device feature of `state`, with the reference's cache protocol (feature.py:24-35): the first device feature asked
    for with a given cache computes all 45 (a miss), the others are served from it (accesses only)."""
    memo = {} if cache is None else cache
    f.accesses += 1
    per_feature = (_DEVICE_VALUES, f.id)
    if per_feature not in memo:
        f.misses += 1
        if _DEVICE_VALUES not in memo:
            memo[_DEVICE_VALUES] = device_features(state, engine)
        vals = memo[_DEVICE_VALUES]
        memo[per_feature] = _typed(f.__name__, vals[f.id], vals[16])
    return memo[per_feature]


def evaluate(state, features, engine=None):
    weighted = isinstance(features, collections.abc.Mapping)
    items = list(features.items()) if weighted else [(f, None) for f in features]
    cache = {}
    out = {}
    for f, w in items:
        v = feature_value(f, state, cache, engine) if isinstance(f, Feature) else f(state, cache=cache)
        out[f] = v * w if weighted else v
    return out
