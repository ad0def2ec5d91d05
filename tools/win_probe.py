#!/usr/bin/env python3
"""K2sW (window-local lockstep kernel) against K2s: kernel ms / placements per s on the BASELINE configs[2] shape, one lane per game,
TB200_NO_WIN=1 = the all-columns K2s.  usage: win_probe.py C spread [C spread ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, random, numpy as np
sys.path.insert(0, %r)
from tetris_ai_b200 import Engine
def seeded_sequence(seed, T):
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]
W6 = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6V = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
C, spread = int(sys.argv[1]), float(sys.argv[2]); G, T = 30, 1000
rng = np.random.default_rng(1)
w = np.array(W6V)[None, :] + rng.normal(0, spread, size=(C, 6))
seqs = np.array([seeded_sequence(100 + g, T) for g in range(G)], dtype=np.uint8)
eng = Engine(0); eng.set_features(W6)
best = 1e9
for _ in range(3):
    r = eng.play_games(w, G, pieces=seqs)
    best = min(best, eng.last_timing()["kernel_ms"])
import hashlib
print("%%.3f %%.4e %%d %%s" %% (best, int(r["results"]["placements"].sum()) / best / 1e-3, eng.last_timing()["lanes"], hashlib.sha256(r["results"].tobytes()).hexdigest()[:12]))
''' % ROOT
def main():
  args = sys.argv[1:] or ["10000", "1.0", "1250", "1.0", "1250", "6.0"]
  for i in range(0, len(args), 2):
      C, spread = args[i], args[i + 1]
      print("C=%s spread=%s" % (C, spread))
      for name, env in (("K2s  lanes 1 (all columns)", {"TB200_LOCK_LANES": "1", "TB200_NO_WIN": "1"}), ("K2sW lanes 1 (window)", {"TB200_LOCK_LANES": "1"}),
                        ("K2s  lanes 4", {"TB200_LOCK_LANES": "4"}), ("auto", {})):
          e = dict(os.environ); e.update(env)
          out = subprocess.run([sys.executable, "-c", CHILD, C, spread], env=e, capture_output=True, text=True, timeout=600).stdout.split()
          print("  %-28s %9s ms  %s placements/s  lanes %s  results %s" % (name, out[0], out[1], out[2], out[3]), flush=True)


if __name__ == "__main__":
    main()
