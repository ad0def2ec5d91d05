#!/usr/bin/env python3
"""K2sP tuning: kernel ms of the per-bucket-width lockstep kernel with and without the pair-of-warps task, for several values of
TB200_LOCK_PAIR (games alive per resident warp from which a bucket is played by pairs of warps; 0 = never) on rank-sized shards of the
BASELINE configs[2] population and on the whole population.  usage: pair_probe.py [C spread ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import win_probe                                    # noqa: E402  (reuses its child script)
args = sys.argv[1:] or ["1250", "1.0", "1250", "6.0", "1800", "1.0", "1800", "6.0", "625", "1.0", "625", "6.0"]
SETTINGS = [("pair off (4 blocks/SM)", {"TB200_LOCK_PAIR": "0"}), ("pair off, 5 blocks/SM", {"TB200_LOCK_PAIR": "0", "TB200_LOCK_MINB": "5"}),
            ("pair 10, 5 blocks/SM (default)", {}), ("pair 10, 4 blocks/SM", {"TB200_LOCK_MINB": "4"}),
            ("pair 7", {"TB200_LOCK_PAIR": "7"}), ("pair 13", {"TB200_LOCK_PAIR": "13"}), ("pair 10, one lane from 32", {"TB200_LOCK_THR": "32,8,2"}),
            ("fixed 1 (K2sW)", {"TB200_LOCK_LANES": "1"})]
for i in range(0, len(args), 2):
    C, spread = args[i], args[i + 1]
    print("C=%s spread=%s" % (C, spread))
    for name, env in SETTINGS:
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, "-c", win_probe.CHILD, C, spread], env=e, capture_output=True, text=True, timeout=600).stdout.split()
        print("  %-32s %9s ms  %s placements/s  lanes %s  results %s" % (name, out[0], out[1], out[2], out[3]) if len(out) >= 4 else "  %-32s FAILED" % name, flush=True)
