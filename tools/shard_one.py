#!/usr/bin/env python3
"""One play of a C x 30 slice of the BASELINE configs[2] population (w6 + N(0, spread), real seeded sequences, T = 1000): the ncu target
for the lockstep kernels on rank-sized shards.  usage: shard_one.py C spread [plays]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import win_probe                                    # noqa: E402
sys.argv = [sys.argv[0]] + (sys.argv[1:3] if len(sys.argv) >= 3 else ["1250", "1.0"])
src = win_probe.CHILD.replace("for _ in range(3):", "for _ in range(%d):" % (int(os.environ.get("PLAYS", "1"))))
exec(compile(src, "shard_one_child", "exec"))
