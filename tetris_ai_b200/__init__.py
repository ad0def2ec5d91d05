"""tetris_ai_b200 — B200-native batched Tetris placement engine.

A from-scratch sm_100a implementation of ONE hot path of CogWorks/Tetris-AI (enumerate placements
-> drop -> clear -> features -> fp64 weighted score -> argmax -> commit, for every game of a
cross-entropy population), behind the reference's own call signatures.  See DESIGN.md.
"""
from .engine import Engine, EngineError  # noqa: F401

__version__ = "0.1"
