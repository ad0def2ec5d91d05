#!/usr/bin/env python3
"""K2sA tuning: the one-lane threshold of STEADY buckets (TB200_LOCK_STEADY = games alive per resident warp from which a bucket that kept
>= 15/16 of the previous stage's games plays with one lane per game; 0 = no special case), and the slot loop unrolled by two
(an experiment build of the library, bin/libtetris_unroll2.so = -DTB_LW_UNROLL=2), on rank-sized shards of the BASELINE configs[2] population
and on the whole population.  usage: steady_probe.py [C spread ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import win_probe                                    # noqa: E402  (reuses its child script)
args = sys.argv[1:] or ["1250", "1.0", "1250", "6.0", "625", "1.0", "625", "6.0", "2500", "1.0", "2500", "6.0", "10000", "1.0", "10000", "6.0"]
U2 = os.path.join(ROOT, "bin", "libtetris_unroll2.so")
SETTINGS = [("steady off", {"TB200_LOCK_STEADY": "0"}), ("steady 6", {"TB200_LOCK_STEADY": "6"}), ("steady 10 (default)", {}), ("steady 14", {"TB200_LOCK_STEADY": "14"}),
            ("fixed 1 (K2sW)", {"TB200_LOCK_LANES": "1"})]
if os.path.exists(U2):
    SETTINGS += [("steady 10, slot loop x2", {"TB200_LIB": U2}), ("fixed 1, slot loop x2", {"TB200_LOCK_LANES": "1", "TB200_LIB": U2})]
for i in range(0, len(args), 2):
    C, spread = args[i], args[i + 1]
    print("C=%s spread=%s" % (C, spread))
    for name, env in SETTINGS:
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "-c", win_probe.CHILD, C, spread], env=e, capture_output=True, text=True, timeout=600).stdout.split()
        print("  %-28s %9s ms  %s placements/s  lanes %s  results %s" % (name, out[0], out[1], out[2], out[3]) if len(out) >= 4 else "  %-28s FAILED" % name, flush=True)
