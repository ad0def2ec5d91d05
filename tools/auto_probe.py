#!/usr/bin/env python3
"""K2sA tuning: kernel ms of the lockstep kernel with per-bucket lane width under different thresholds (TB200_LOCK_THR=a,b,c = games alive per
resident warp from which a bucket plays with 1 / 4 / 8 lanes per game) and resident blocks per SM (TB200_LOCK_MINB), against the fixed widths.
usage: auto_probe.py [C spread ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import win_probe                                    # noqa: E402  (reuses its child script)
args = sys.argv[1:] or ["1250", "1.0", "1250", "6.0", "3000", "1.0", "3000", "6.0", "10000", "6.0", "10000", "1.0"]
SETTINGS = [("fixed 1 (K2sW)", {"TB200_LOCK_LANES": "1"}), ("fixed 4", {"TB200_LOCK_LANES": "4"}), ("fixed 8", {"TB200_LOCK_LANES": "8"}),
            ("auto (default)", {}), ("auto 24,8,2", {"TB200_LOCK_THR": "24,8,2"}), ("auto 12,6,2", {"TB200_LOCK_THR": "12,6,2"}),
            ("auto 12,8,2", {"TB200_LOCK_THR": "12,8,2"}), ("auto 16,8,2", {"TB200_LOCK_THR": "16,8,2"})]
for i in range(0, len(args), 2):
    C, spread = args[i], args[i + 1]
    print("C=%s spread=%s" % (C, spread))
    for name, env in SETTINGS:
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "-c", win_probe.CHILD, C, spread], env=e, capture_output=True, text=True, timeout=600).stdout.split()
        print("  %-24s %9s ms  %s placements/s  lanes %s  results %s" % (name, out[0], out[1], out[2], out[3]), flush=True)
