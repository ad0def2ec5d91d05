#!/usr/bin/env python3
"""bench.py — placements evaluated / s (and CE generations / s) of the GPU placement engine.

Contract (see the task statement):  python bench.py --gpus N --steps K --warmup W [--impl reference]
prints ONE JSON line from rank 0.

Headline workload (BASELINE.json configs[1]): one cross-entropy generation of 100 candidates x 10 games,
one-piece placement search, the 6-feature "w6" set, injected sequences random.Random(1000..1009),
cap T = 1000 pieces.  The population is generation 1 of the canonical CE run (random.Random(42),
sigma = 10; mean/std after generation 0 from tests/golden/ce.json): 6 680 307 placements in
294 598 pieces — a mid-training generation in which many games reach the cap.  A step = one pass of
the hot path over that population = evaluating the whole generation (the play kernels + the fitness
reduction).  For N > 1 (weak scaling) the population is N replicas of that generation, one
100-candidate shard per rank (identical work on every rank), and the per-candidate fitness of all
100*N candidates is all-gathered over NCCL inside the step.

value      : device-timed (CUDA events on the launching stream), inputs resident in HBM.
e2e        : the same step through the public API with HOST buffers in pinned memory, copies inside the timed region, wall
             clock: N = 1 PopulationEvaluator.evaluate -> tb200_play_games (host mode); N > 1 ShardedEvaluator.evaluate
             (H2D weights + pieces, device-mode play, NCCL all-gather of the device fitness, one D2H of f64[C]).
roofline   : this path is integer/bitwise work, so the bound is the SM integer-issue rate, not HBM
             or tensor cores (documented deviation from the "hbm"|"tensor" enum): achieved =
             placements/s x 220 algorithmic thread-ops per placement against the
             INT_PEAK measured IN THIS RUN by bin/int_peak (the committed profiles/INT_PEAK_r01.json is shown beside it).
population_sharded : BASELINE configs[2] — the north-star multi-GPU split — at EVERY N (1 included): one generation of
             10 000 candidates x 30 games (T = 1000) sharded by candidate over the N ranks through ShardedEvaluator
             (strong scaling: total work fixed), for a converged population (w6 + N(0,1)) and for generation 0 of the
             canonical run, whose refit digest must equal the C-oracle fixture tests/golden/ce10k_gen0.json.gz.
cpu_baseline: the reference's CPU path on all host cores over a bounded sample of the headline population: the UNMODIFIED
             reference (cogworks.*, pip-installed under baseline/_ref — kind "reference") when it is importable, else
             oracle/pyoracle.py (numpy port, kind "port").  --impl reference times the same thing as its own arm.
"""
import argparse
import hashlib
import heapq
import json
import os
import random
import socket
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W6_NAMES = ["landing_height", "eroded_cells", "row_trans", "col_trans", "pits", "cuml_wells"]
W6_VALS = [-4.500158825082766, 3.4181268101392694, -3.2178882868487753, -9.348695305445199, -7.899265427351652, -3.3855972247263626]
W6_OPS_PER_PLACEMENT = 220.0          # SURVEY.md §8d algorithmic thread-ops per placement, 6-feature set
CANDIDATES_PER_GPU, GAMES, T_CAP = 100, 10, 1000
METRIC, UNIT = "placements evaluated/sec (CE generation, 100 candidates x 10 games per GPU)", "placements/s"


# ---------------------------------------------------------------------------------------------
# stdout hygiene: the contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner and, with
# NCCL_DEBUG=INFO, its topology lines) write to fd 1, so fd 1 points at stderr while the benchmark runs and the JSON line
# goes to the saved descriptor.
# ---------------------------------------------------------------------------------------------
_REAL_STDOUT = None




def seeded_sequence(seed, T):
    """The canonical injected piece sequence: ONE random.Random(seed), T draws of randrange(7)."""
    r = random.Random(seed)
    return [r.randrange(7) for _ in range(T)]


def guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
def canonical_population(n_total):
    """Generation-1 population of the canonical CE run, extended to n_total candidates for N > 1."""
    ce = json.load(open(os.path.join(ROOT, "tests", "golden", "ce.json")))["ce"]
    rng = random.Random(ce["rng_seed"])
    for _ in range(ce["width"]):                       # generation 0 draws (consumed, not used)
        [rng.gauss(0, ce["stdev"]) for _ in range(6)]
    mean = [float.fromhex(x) for x in ce["generations"][0]["mean"]]
    std = [float.fromhex(x) for x in ce["generations"][0]["std"]]
    pop = [[rng.gauss(mean[i], std[i]) for i in range(6)] for _ in range(n_total)]
    seqs = []
    for g in range(GAMES):
        r = random.Random(ce["seed0"] + g)
        seqs.append([r.randrange(7) for _ in range(T_CAP)])
    return np.array(pop, dtype=np.float64), np.array(seqs, dtype=np.uint8)


def workload_config(n_gpus):
    return {
        "workload": "BASELINE configs[1]: CE generation, 100 candidates x 10 games per GPU, lookahead 1, w6 features, T=1000, "
                    "population = generation 1 of the canonical run (Random(42), sigma 10)",
        "candidates_per_gpu": CANDIDATES_PER_GPU, "games_per_candidate": GAMES, "cap_pieces": T_CAP, "features": W6_NAMES,
        "lookahead": 1, "parallelism": "candidates sharded per rank (population = N replicas of the canonical generation); fitness all-gather (NCCL)" if n_gpus > 1 else "single GPU",
        "l2": "L2 flushed (256 MiB write) between timed steps, outside the per-step event pairs",
    }


# ---------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference when importable (baseline/_ref), else the numpy port — also used for cpu_baseline
# ---------------------------------------------------------------------------------------------
_CPU_JOB = None
_REF = None


def load_reference():
    """The unmodified reference package (pip-installed by __graft_entry__.build() into the git-ignored baseline/_ref, which
    travels to the GPU box), or None.  Nothing here reads /root/reference."""
    global _REF
    if _REF is not None:
        return _REF or None
    path = os.path.join(ROOT, "baseline", "_ref")
    _REF = False
    if os.path.isdir(os.path.join(path, "cogworks")):
        import collections
        import collections.abc
        collections.Mapping = collections.abc.Mapping          # removed in Python 3.10; feature.py:52, learning.py:9 use it
        sys.path.insert(0, path)
        try:
            from cogworks.tetris import features as rf, simulator as rs
            from cogworks.tetris.game import Board, State, zoids
            _REF = {"features": rf, "sim": rs, "Board": Board, "State": State, "pieces": list(zoids.classic)}
        except Exception as e:                                   # pragma: no cover
            sys.stderr.write("bench.py: baseline/_ref present but not importable (%s); using the numpy port\n" % e)
        finally:
            sys.path.remove(path)
    return _REF or None


def _ref_one(i):
    """One candidate through the reference's own simulate / move_drop / policy_best / features (canonical harness of
    SURVEY.md §8c).  Returns (total lines, scorer calls = placements evaluated)."""
    ref = load_reference()
    _fids, w, seqs = _CPU_JOB
    items = [(getattr(ref["features"], n), float(x)) for n, x in zip(W6_NAMES, w[i])]
    calls = [0]

    def scorer(state):
        calls[0] += 1
        cache = {}
        acc = 0.0
        for f, wt in items:
            acc += float(f(state, cache)) * wt
        return acc
    policy = ref["sim"].policy_best(scorer, lambda maxes: maxes[0])
    total = 0
    for s in seqs:
        last = ref["State"](None, ref["Board"](20, 10))
        for last in ref["sim"].simulate(last, iter([ref["pieces"][int(x)] for x in s]), ref["sim"].move_drop, policy, 1):
            pass
        total += last.lines_cleared()
    return total, calls[0]


def _port_one(i):
    from oracle import pyoracle
    fids, w, seqs = _CPU_JOB
    total = 0
    for s in seqs:
        total += pyoracle.play_game(fids, list(w[i]), [int(x) for x in s])["lines"]
    return total, 0


def cpu_rate(pop, seqs, n_cand, n_games, cores):
    """placements/s of the reference's CPU path on `cores` processes over the first n_cand candidates x n_games games.
    -> (rate, placements, seconds, kind)."""
    import multiprocessing as mp
    from oracle import coracle, pyoracle
    global _CPU_JOB
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    sub = seqs[:n_games]
    _CPU_JOB = (fids, pop, sub)
    kind = "reference" if load_reference() else "port"
    placements = None
    if kind == "port":       # the port does not count scorer calls: count them with the C oracle (same games, bit-identical by the parity tests), before t0
        placements = int(coracle.play_games(fids, pop[:n_cand], n_games, pieces=sub, threads=cores)["placements"].sum())
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        out = pool.map(_ref_one if kind == "reference" else _port_one, range(n_cand), chunksize=1)
    dt = time.perf_counter() - t0
    if kind == "reference":
        placements = sum(c for _, c in out)
    return placements / dt, placements, dt, kind


def cpu_c_rate(pop, seqs, cores):
    from oracle import coracle, pyoracle
    fids = [pyoracle.FEATURE_ID[n] for n in W6_NAMES]
    t0 = time.perf_counter()
    res = coracle.play_games(fids, pop, GAMES, pieces=seqs, threads=cores)
    dt = time.perf_counter() - t0
    return int(res["placements"].sum()) / dt, dt


def pick_cpu_sample(cores):
    # ONE sample definition for both CPU legs (the cpu_baseline object of the GPU arm and every step of --impl reference):
    # the first 2 candidates per core x the first 5 games — ~12 k placements/s/core for the reference, ~18 k for the port,
    # i.e. roughly 5 - 10 s per pass on the GPU box's host.
    return min(CANDIDATES_PER_GPU, 2 * cores), 5


def cpu_sample_text(kind, n_cand, n_games, placements, dt):
    what = ("UNMODIFIED reference (cogworks.tetris.simulator.simulate / move_drop / policy_best + cogworks.tetris.features, from baseline/_ref)"
            if kind == "reference" else "oracle/pyoracle.py (numpy port of the reference; baseline/_ref not importable here)")
    return "%s, fork pool, one candidate per task: first %d candidates x first %d games of this population: %d placements in %.1f s" % (
        what, n_cand, n_games, placements, dt)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pop, seqs = canonical_population(CANDIDATES_PER_GPU)
    n_cand, n_games = pick_cpu_sample(cores)
    rates, times = [], []
    kind, placements = "port", 0
    for i in range(args.warmup + args.steps):
        rate, placements, dt, kind = cpu_rate(pop, seqs, n_cand, n_games, cores)
        if i >= args.warmup:
            rates.append(rate)
            times.append(dt)
    value = statistics.mean(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 bitboards + f64 score", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": cpu_sample_text(kind, n_cand, n_games, placements, statistics.mean(times))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "each step = one pass over the stated sample on all host cores (learning.py:38 granularity: one candidate per task)",
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML, every ~2 ms, in a thread)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.stop_flag, self.thread, self.max_mhz = index, [], set(), False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measure_int_peak(local):
    """Run the integer-issue micro-benchmark (csrc/int_peak.cu -> bin/int_peak, ~0.2 s) on this rank's GPU, now."""
    exe = os.path.join(ROOT, "bin", "int_peak")
    if not os.path.exists(exe):
        return None
    out = os.path.join("/tmp", "int_peak_%d.json" % os.getpid())
    env = dict(os.environ)
    vis = env.get("CUDA_VISIBLE_DEVICES")
    env["CUDA_VISIBLE_DEVICES"] = vis.split(",")[local] if vis else str(local)
    try:
        subprocess.run([exe, out], env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=120)
        return json.load(open(out))
    except Exception:
        return None
    finally:
        if os.path.exists(out):
            os.remove(out)


# ---------------------------------------------------------------------------------------------
# BASELINE configs[2]: the north-star multi-GPU split, at every N
# ---------------------------------------------------------------------------------------------
def refit_digest(pop, fitness, keep):
    """learning.py:41-46 on the host: elite selection + mean / std refit; digest of the refitted doubles."""
    scored = [(fitness[i], i) for i in range(len(pop))]
    elite = [pop[i] for _, i in heapq.nlargest(keep, scored)]
    mean = [np.mean([e[d] for e in elite]) for d in range(6)]
    std = [np.std([e[d] for e in elite]) for d in range(6)]
    return hashlib.sha256(("".join(float(v).hex() for v in mean + std)).encode()).hexdigest()[:16]


def population_sharded(eng, dev, world, int_peak, which, reps=3):
    import gzip
    import torch
    import torch.distributed as dist
    from tetris_ai_b200.learning import PopulationEvaluator, ShardedEvaluator, shard_bounds
    C, G, T, keep = 10000, 30, 1000, 1000
    fixture = None
    if which == "gen0":
        fx = os.path.join(ROOT, "tests", "golden", "ce10k_gen0.json.gz")
        fixture = json.load(gzip.open(fx, "rt")) if os.path.exists(fx) else None
        rng = random.Random(42)
        pop = [[rng.gauss(0, 10.0) for _ in range(6)] for _ in range(C)]
        seqs = np.array([seeded_sequence(1000 + g, T) for g in range(G)], dtype=np.uint8)
        label = "generation 0 of the canonical configs[2] run (Random(42), sigma 10, seeds 1000..1029): a FRESH population, most games die within a few dozen pieces"
    else:
        rng = np.random.default_rng(1)
        pop = (np.array(W6_VALS)[None, :] + rng.normal(0, 1.0, size=(C, 6))).tolist()
        seqs = np.array([seeded_sequence(100 + g, T) for g in range(G)], dtype=np.uint8)
        label = "w6 + N(0,1) noise: a CONVERGED population, most games reach the cap (the population_regime population)"
    w = np.array(pop, dtype=np.float64)
    local_ev = PopulationEvaluator(eng, W6_NAMES, G, pieces=seqs)
    ev = ShardedEvaluator(local_ev, device=dev)
    ev.collect_timing = True
    ev.collect_placements = True
    rows, digest, placements = [], None, 0
    for i in range(reps + 1):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fit = ev.evaluate(w)
        t1 = time.perf_counter()
        digest = refit_digest(pop, fit.tolist(), keep)
        t2 = time.perf_counter()
        tm = dict(ev.last_timing)
        mine = torch.tensor([tm["evaluate_ms"], tm["allgather_ms"], tm["d2h_ms"], tm["host_prep_ms"], 1e3 * (t1 - t0), 1e3 * (t2 - t1), float(tm["placements"])],
                            dtype=torch.float64, device=dev)
        allr = torch.empty(world * mine.numel(), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.cpu().numpy().reshape(world, -1)
        if i > 0:                      # first pass = warm-up (buffer allocation, workspace growth, NCCL channel set-up)
            rows.append(allr)
        placements = int(allr[:, 6].sum())
    r = np.mean(np.stack(rows), axis=0)                    # [world][7], mean over the timed repetitions
    evaluate_ms = float(r[:, 0].max())                     # the slowest rank's play kernels (+ H2D): what the generation waits for
    total_ms = float(r[:, 4].max())
    out = {
        "workload": "BASELINE configs[2]: 10 000 candidates x 30 games, T = 1000, w6, lookahead 1, sharded by candidate over %d rank(s) "
                    "(learning.ShardedEvaluator: H2D weights, device-mode play of the local slice, ONE NCCL all-gather of f64[C], one D2H); %s" % (world, label),
        "scaling": "strong", "n_gpus": world, "repetitions": reps,
        "placements": placements, "candidates_per_rank": [b - a for a, b in (shard_bounds(C, world, k) for k in range(world))],
        "evaluate_ms": evaluate_ms, "evaluate_ms_per_rank": [float(x) for x in r[:, 0]],
        "allgather_ms": float(r[:, 1].min()), "allgather_ms_per_rank": [float(x) for x in r[:, 1]],
        "d2h_ms": float(r[:, 2].max()), "host_prep_ms": float(r[:, 3].max()),
        "shard_call_ms": total_ms,                         # ShardedEvaluator.evaluate wall clock, slowest rank
        "host_ms": float(r[:, 5].max()),                   # elite selection + mean / std refit (learning.py:41-46), pure host
        "gen_ms": total_ms + float(r[:, 5].max()),
        "placements_per_s_evaluate": placements / (evaluate_ms * 1e-3), "placements_per_s_gen": placements / ((total_ms + float(r[:, 5].max())) * 1e-3),
        "roofline_frac_int_issue": placements / (evaluate_ms * 1e-3) * W6_OPS_PER_PLACEMENT / (int_peak * world),
        "refit_digest": digest,
        "note": "allgather_ms = the fastest rank's figure (the others include waiting for the slowest rank's kernels); "
                "host_ms excludes the reference driver's own rng.gauss / test_f loop (learning.py:32-38), which this object does not run",
    }
    if fixture is not None:
        out["refit_digest_oracle_fixture"] = fixture["refit_digest"]
        out["digest_matches_oracle_fixture"] = bool(digest == fixture["refit_digest"])
        out["placements_match_oracle_fixture"] = bool(placements == fixture["placements"])
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from tetris_ai_b200 import Engine, _lib
    from tetris_ai_b200.learning import PopulationEvaluator, ShardedEvaluator

    guard_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL at every N (population_sharded runs at N = 1 too); INFO so that the ranks / transports are visible in the log
    # (fd 1 already points at stderr, so stdout stays the one JSON line)
    os.environ["NCCL_DEBUG"] = os.environ.get("TB_NCCL_DEBUG", "INFO")
    if world == 1 and "MASTER_ADDR" not in os.environ:
        os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(free_port()), "RANK": "0", "WORLD_SIZE": "1"})
    dist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)

    # weak scaling: the global population is `world` replicas of the canonical 100-candidate generation, one replica
    # per rank, so that every rank has exactly the same work and the N-GPU figure isolates launch skew + the all-gather
    pop, seqs = canonical_population(CANDIDATES_PER_GPU)
    eng = Engine(local)
    eng.set_features(W6_NAMES)
    C, G, n_games = CANDIDATES_PER_GPU, GAMES, CANDIDATES_PER_GPU * GAMES

    # ---- device-resident buffers (value) ----
    d_w = torch.from_numpy(pop).to(dev)
    d_p = torch.from_numpy(seqs).to(dev)
    d_res = torch.zeros(n_games * _lib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_fit = torch.zeros(C, dtype=torch.float64, device=dev)
    d_all = torch.zeros(C * world, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step_device():
        eng.play_games_device(stream.cuda_stream, C, G, d_w.data_ptr(), d_res.data_ptr(), T_CAP, pieces_ptr=d_p.data_ptr(),
                              n_sequences=G, fitness_ptr=d_fit.data_ptr(), lanes=args.lanes)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_fit)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)                                    # L2 flush, outside the event pair
        a.record(stream)
        step_device()
        b.record(stream)
    sync_all()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    # dominant kernel(s) alone: the library's own CUDA-event pair on the launching stream around the play kernels,
    # read back after each of K further steps; average launch duration
    kms = []
    for _ in range(args.steps):
        flush.fill_(1)
        step_device()
        kms.append(eng.last_timing()["kernel_ms"])
    sync_all()
    eng.check_status()
    kernel_ms = sum(kms) / len(kms)
    lanes_used = eng.last_timing()["lanes"]
    launches_per_step = int(eng.last_timing()["kernels_launched"])      # play kernels (1, or the re-spreading stages) + fitness
    res = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=_lib.RESULT_DTYPE)
    placements = int(res["placements"].sum())
    pieces = int(res["pieces"].sum())

    # ---- e2e: public API, pinned host buffers, copies inside the timed region ----
    h_w = torch.from_numpy(np.tile(pop, (world, 1))).pin_memory()
    evalr = PopulationEvaluator(eng, W6_NAMES, G, pieces=seqs, lanes=args.lanes)
    if world == 1:
        e2e_call, e2e_api = (lambda: evalr.evaluate(h_w.numpy())), "PopulationEvaluator.evaluate -> tb200_play_games (host buffers, pinned)"
    else:
        shard = ShardedEvaluator(evalr, device=dev)
        e2e_call = lambda: shard.evaluate(h_w.numpy())      # noqa: E731
        e2e_api = ("ShardedEvaluator.evaluate: H2D of the rank's weights + the piece sequences (pinned), tb200_play_games (device mode), NCCL all-gather "
                   "of the device fitness, one D2H of f64[C] (pinned)")
    for _ in range(3):
        e2e_call()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_call()
    sync_all()
    e2e_s = time.perf_counter() - t0
    if world == 1:
        tm = eng.last_timing()
        h2d_b, d2h_b = int(tm["h2d_bytes"]), int(tm["d2h_bytes"])
    else:
        h2d_b, d2h_b = int(shard.last_bytes["h2d"]), int(shard.last_bytes["d2h"])
    clocks = sampler.stop() if rank == 0 else None

    # ---- reduce over ranks: max time, sum placements ----
    stats = torch.tensor([total_ms, e2e_s, float(placements), float(pieces), kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms, e2e_s, kernel_ms = float(mx[0]), float(mx[1]), float(mx[4])
        placements_all, pieces_all = int(sm[2]), int(sm[3])
    else:
        placements_all, pieces_all = placements, pieces

    # ---- integer-issue peak, measured now on this GPU (rank 0), next to the committed figure ----
    peak_live = measure_int_peak(local) if rank == 0 else None
    peaks_path = os.path.join(ROOT, "profiles", "INT_PEAK_r01.json")
    peak_committed = json.load(open(peaks_path))["int_peak_thread_ops_per_s"] if os.path.exists(peaks_path) else None
    int_peak = (peak_live or {}).get("int_peak_thread_ops_per_s") or peak_committed or 148 * 128 * 1.965e9
    pk = torch.tensor([int_peak], dtype=torch.float64, device=dev)
    dist.broadcast(pk, 0)
    int_peak = float(pk.item())

    # ---- BASELINE configs[2] through ShardedEvaluator, every N ----
    sharded = {}
    if not args.no_sharded:
        for which in ("converged", "gen0"):
            sharded[which] = population_sharded(eng, dev, world, int_peak, which)
    if rank != 0:
        dist.destroy_process_group()
        return

    value = placements_all * args.steps / (total_ms * 1e-3)
    e2e_value = placements_all * args.steps / e2e_s
    # roofline of the dominant kernel(s) (this rank's launches): algorithmic thread-ops / kernel time
    k_rate = placements / (kernel_ms * 1e-3)
    achieved = k_rate * W6_OPS_PER_PLACEMENT
    hbm_peak = 6554.2
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    alg_bytes = C * 6 * 8 + seqs.size + n_games * _lib.RESULT_DTYPE.itemsize + C * 8
    traffic, traffic_src = None, None
    for name in ("ncu_r02_staged_summary.json", "ncu_r01_staged_summary.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
                traffic_src = "profiles/%s (one ncu --set full capture of this command, summed over the play launches of one step; NOT measured in this run)" % name
            except Exception:
                traffic = None
            if traffic is not None:
                break
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 bitboards + f64 score", "data": "synthetic", "config": workload_config(world),
        "generations_per_s": args.steps / (total_ms * 1e-3),
        "placements_per_step": placements_all, "pieces_per_step": pieces_all, "lanes_per_game": lanes_used,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                "ms_per_step": 1e3 * e2e_s / args.steps, "api": e2e_api},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "int-issue", "achieved": achieved / 1e12, "peak": int_peak / 1e12, "unit": "Tops/s (thread integer ops)",
                     "frac": achieved / int_peak, "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": ("the re-spreading stages of one play (%d stream-ordered launches: play_games_w6fast_kernel<staged> x2, play_games_w6pair_kernel), "
                                "one event pair around all of them" % (launches_per_step - 1)
                                if launches_per_step > 2 else
                                {128: "play_games_w6quad_kernel", 64: "play_games_w6pair_kernel", 32: "play_games_w6fast_kernel"}.get(lanes_used, "play_games_w6grp_kernel<%d>" % lanes_used)),
                     "kernel_ms": kernel_ms, "kernel_launches_averaged": len(kms),
                     "note": "1000 games on 592 warp schedulers: the run time of this workload is the serial chain of its longest game "
                             "(1000 pieces x the per-piece latency), so the fraction is low by construction; see population_sharded / "
                             "population_regime / throughput_regime for the same evaluation code with the machine full",
                     "algorithmic_ops_per_placement": W6_OPS_PER_PLACEMENT,
                     "peak_source": ("measured in this run: bin/int_peak on this GPU" if peak_live else "profiles/INT_PEAK_r01.json (bin/int_peak did not run here)"),
                     "peak_committed_r01": (peak_committed / 1e12) if peak_committed else None,
                     "peak_ops_detail": {k: v["thread_ops_per_clk_per_sm"] for k, v in (peak_live or {}).get("ops", {}).items()} or None,
                     "hbm": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9,
                             "peak_gbs": hbm_peak, "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak}},
        "clocks": clocks,
        "wall_s_timed_region": t_wall,
        "play_launches_per_step": launches_per_step - 1,
    }
    if sharded:
        line["population_sharded"] = sharded["converged"]
        line["population_sharded_gen0"] = sharded["gen0"]
    if world == 1 and not args.no_sweep:
        line["throughput_regime"] = throughput_regime(eng, int_peak, n=262144)
        line["population_regime"] = population_regime(eng, int_peak)
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        n_cand, n_games_s = pick_cpu_sample(cores)
        rate, pl, dt, kind = cpu_rate(pop, seqs, n_cand, n_games_s, cores)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": cpu_sample_text(kind, n_cand, n_games_s, pl, dt)}
        c_rate, c_dt = cpu_c_rate(pop, seqs, cores)
        line["cpu_baseline_c"] = {"value": c_rate, "unit": UNIT, "cores": cores, "kind": "port",
                                  "sample": "oracle/coracle.c (plain-C restatement, pthreads), the whole step (%d placements) in %.2f s" % (placements, c_dt)}
    emit(line)
    dist.destroy_process_group()


def throughput_regime(eng, int_peak, n=65536, T=400):
    """BASELINE configs[4] shape (not the headline): n concurrent games, each its own candidate (w6 + N(0, 3) noise) and its
    own counter-based piece stream; kernel-only timing.  Shows the kernel when the machine is full."""
    w6 = np.array(W6_VALS)
    rng = np.random.default_rng(1)
    w = w6[None, :] + rng.normal(0, 3.0, size=(n, 6))
    best, out = 1e9, None
    for _ in range(3):
        out = eng.play_games(w, 1, T=T, seed=5, seq_of_game=np.arange(n, dtype=np.int32))
        best = min(best, eng.last_timing()["kernel_ms"])
    pl = int(out["results"]["placements"].sum())
    rate = pl / (best * 1e-3)
    return {"workload": "%d concurrent games, cap %d pieces, one candidate and one hashed piece stream per game, w6, lookahead 1" % (n, T),
            "placements_per_s": rate, "kernel_ms": best, "placements": pl, "lanes_per_game": eng.last_timing()["lanes"],
            "roofline_frac_int_issue": rate * W6_OPS_PER_PLACEMENT / int_peak}


def population_regime(eng, int_peak, C=10000, G=30, T=1000):
    """BASELINE configs[2] shape on ONE GPU through the host-mode C ABI (kernel-only timing): C candidates x G games on G shared
    seeded piece sequences, weights = w6 + N(0, 1) noise (a converged CE population: most games reach the cap)."""
    w6 = np.array(W6_VALS)
    rng = np.random.default_rng(1)
    w = w6[None, :] + rng.normal(0, 1.0, size=(C, 6))
    seqs = np.array([seeded_sequence(100 + g, T) for g in range(G)], dtype=np.uint8)
    best, out = 1e9, None
    for _ in range(2):
        out = eng.play_games(w, G, pieces=seqs)
        best = min(best, eng.last_timing()["kernel_ms"])
    pl = int(out["results"]["placements"].sum())
    rate = pl / (best * 1e-3)
    return {"workload": "%d candidates x %d games on %d shared piece sequences, cap %d pieces, w6, lookahead 1 (the whole BASELINE configs[2] population on one GPU)" % (C, G, G, T),
            "placements_per_s": rate, "kernel_ms": best, "placements": pl, "lanes_per_game": eng.last_timing()["lanes"],
            "roofline_frac_int_issue": rate * W6_OPS_PER_PLACEMENT / int_peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the throughput-regime extras")
    ap.add_argument("--no-sharded", action="store_true", help="skip the population_sharded objects")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
