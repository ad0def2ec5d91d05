"""Feature handles — host-side mirror of cogworks/tetris/features.py.

In the reference each feature is a Python callable feature(state, cache=None) (feature.py:24) that
is also used as a dict key in the weights mapping (learning.py:38).  Here the 45 numeric features
are handles with the same names (module attributes, `__name__`) and the ids the kernels use
(= source order of the reference file); the arithmetic itself runs on the GPU
(csrc/tetris_core.cuh: eval_features).  Calling a handle on a State evaluates that feature on the
device through Engine.eval_features.
"""
from __future__ import annotations

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
# features whose reference value is a Python float (true division somewhere); the rest are ints
FLOAT_FEATURES = frozenset(["mean_ht", "max_ht_diff", "min_ht_diff", "cuml_cleared", "cuml_wells", "mean_pit_depth",
                            "d_mean_ht", "cuml_eroded"])


class Feature:
    """Hashable handle usable as a weights-mapping key, like the reference's feature functions.

    accesses / misses mirror the counters of the reference's memo wrapper (feature.py:27,31,36-37): `accesses` counts every
    evaluation request for this feature, `misses` the requests that had to compute it — here, the requests that were not
    served from the per-evaluation cache of device values (one K3 launch fills all 45 at once)."""
    __slots__ = ("__name__", "id", "accesses", "misses")

    def __init__(self, name, fid):
        self.__name__ = name
        self.id = fid
        self.accesses = 0
        self.misses = 0

    def __call__(self, state, cache=None):
        from ..feature import feature_value
        return feature_value(self, state, cache)

    def __repr__(self):
        return "<feature %s #%d>" % (self.__name__, self.id)


def is_helper(feature):
    """cogworks/tetris/features.py:8 — helper features start with '__'; none of the 45 handles is one."""
    return feature.__name__.startswith("__")


ALL = tuple(Feature(n, i) for i, n in enumerate(FEATURE_NAMES))
globals().update({f.__name__: f for f in ALL})

__all__ = ["FEATURE_NAMES", "FEATURE_ID", "Feature", "ALL", "is_helper"] + list(FEATURE_NAMES)
