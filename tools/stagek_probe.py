#!/usr/bin/env python3
"""How much of a single-wave shard's time is the stage boundary (every warp waits there for the slowest task of its sequence)?
Kernel ms of the one-lane lockstep kernel for stage lengths TB200_LOCK_K = 64 (default) .. 500 on rank-sized shards.
usage: stagek_probe.py [C spread ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import win_probe                                    # noqa: E402  (reuses its child script)
args = sys.argv[1:] or ["1250", "1.0", "2500", "1.0", "10000", "1.0"]
SETTINGS = [("fixed 1, K %s" % k, {"TB200_LOCK_LANES": "1", "TB200_LOCK_K": k}) for k in ("64", "128", "256", "500")] + \
           [("auto, K %s" % k, {"TB200_LOCK_K": k}) for k in ("64", "256")]
for i in range(0, len(args), 2):
    C, spread = args[i], args[i + 1]
    print("C=%s spread=%s" % (C, spread))
    for name, env in SETTINGS:
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "-c", win_probe.CHILD, C, spread], env=e, capture_output=True, text=True, timeout=600).stdout.split()
        print("  %-28s %9s ms  %s placements/s  lanes %s  results %s" % (name, out[0], out[1], out[2], out[3]) if len(out) >= 4 else "  %-28s FAILED" % name, flush=True)
