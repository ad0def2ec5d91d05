import sys, cProfile, pstats, io, time
sys.path.insert(0, "/root/repo")
import numpy as np, bench
from tetris_ai_b200 import Engine
from tetris_ai_b200.learning import PopulationEvaluator
pop, seqs = bench.canonical_population(100)
eng = Engine(0)
ev = PopulationEvaluator(eng, bench.W6_NAMES, 10, pieces=seqs)
for _ in range(5): ev.evaluate(pop)
t0 = time.perf_counter()
for _ in range(500): ev.evaluate(pop)
print("ms per evaluate", (time.perf_counter() - t0) / 500 * 1e3, "last total_ms", eng.last_timing()["total_ms"])
pr = cProfile.Profile(); pr.enable()
for _ in range(500): ev.evaluate(pop)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
