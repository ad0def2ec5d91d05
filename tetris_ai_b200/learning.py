"""Cross-entropy driver with the reference's signature (cogworks/learning.py:7) plus the adapters
that route its candidate evaluation to the GPU.

cross_entropy() itself stays on the host, exactly as in the reference: sample `width` weight
vectors, evaluate them through map_f / test_f, keep the `keep` best, refit mean and standard
deviation.  The evaluation — >99.9 % of the reference's run time — is what moves to the device:
PopulationEvaluator.test_f only records the candidate's weights and PopulationEvaluator.map_f
plays the whole generation in ONE kernel launch, so the unmodified reference driver
(cogworks.learning.cross_entropy) can be used just as well as this mirror.

Multi-GPU: ShardedEvaluator gives every rank a contiguous slice of the candidates and all-gathers
the per-candidate fitness (f64[C]) — the only data-path collective.  Every rank
runs the same seeded driver, so after the all-gather all ranks refit identical distributions.
"""
from __future__ import annotations

import collections
import collections.abc
import heapq

import numpy as np


def cross_entropy(feats, stdev, width, keep, rng, test_f, noise_f=None, map_f=map):
    """Infinite generator of (OrderedDict feature->mean, OrderedDict feature->stdev), one per generation."""
    if isinstance(feats, collections.abc.Mapping):
        keys = list(feats.keys())
        mu = [feats[k] for k in keys]
    else:
        keys = list(feats)
        mu = [0] * len(keys)
    sigma = [stdev[k] for k in keys] if isinstance(stdev, collections.abc.Mapping) else [stdev] * len(keys)
    dims = range(len(keys))
    while True:
        if noise_f is not None:
            sigma = [noise_f(s) for s in sigma]
        population = []
        for _ in range(width):                    # candidate-major, feature-minor draw order
            population.append([rng.gauss(mu[d], sigma[d]) for d in dims])
        scored = map_f(lambda i: (test_f(dict(zip(keys, population[i]))), i), range(width))
        elite = [population[i] for _, i in heapq.nlargest(keep, scored)]
        for d in dims:
            column = [cand[d] for cand in elite]
            mu[d] = np.mean(column)
            sigma[d] = np.std(column)
        yield collections.OrderedDict(zip(keys, mu)), collections.OrderedDict(zip(keys, sigma))


class _Pending:
    """Placeholder returned by test_f until map_f has run the batch."""
    __slots__ = ("slot",)

    def __init__(self, slot):
        self.slot = slot


class PopulationEvaluator:
    """test_f / map_f pair that evaluates a whole generation in one device launch.

    fitness of a candidate = total lines cleared over its `games` games / games  (SURVEY.md §8c).
    pieces: u8[games][T] injected sequences, or None for the counter-based stream (seed, game index, t).
    """

    def __init__(self, engine, features, games, pieces=None, T=None, seed=0, lookahead=1, lanes=0):
        self.engine = engine
        self.features = list(features)
        self.names = [getattr(f, "__name__", f) for f in self.features]
        self.games = int(games)
        self.pieces = None if pieces is None else np.ascontiguousarray(pieces, dtype=np.uint8)
        self.T = int(T if T is not None else self.pieces.shape[1])
        self.seed, self.lookahead, self.lanes = int(seed), int(lookahead), int(lanes)
        self._batch = []
        self._keys_ok = None                  # key tuple of the last validated weights mapping
        self._bufs = None                     # page-locked staging buffers, reused every generation (-> CUDA-graph replay)
        self.last = None                      # raw device results of the last generation
        self.generations = 0

    # -- the two callables handed to cross_entropy ---------------------------------------------
    def test_f(self, weights):
        # called once per candidate (learning.py:38), 10 000 times per generation in BASELINE configs[2]: the key order is
        # validated by name the first time a key tuple is seen and by identity afterwards
        keys = tuple(weights)
        if keys != self._keys_ok:
            names = [getattr(k, "__name__", k) for k in keys]
            if names != self.names:
                raise ValueError("weights mapping keys %r do not match the evaluator's feature order %r" % (names, self.names))
            self._keys_ok = keys
        self._batch.append(list(weights.values()))
        return _Pending(len(self._batch) - 1)

    def map_f(self, fn, indices):
        self._batch = []
        tagged = [fn(i) for i in indices]                 # cheap: only records weights
        fitness = self.evaluate(np.array(self._batch, dtype=np.float64)).tolist()      # Python floats, same bits
        return [(fitness[p.slot], i) for p, i in tagged]

    # -- batch evaluation ----------------------------------------------------------------------------
    def evaluate(self, weights):
        """weights f64[C][F] -> fitness f64[C]."""
        self.engine.set_features(self.names)
        weights = np.asarray(weights, dtype=np.float64)
        C = weights.shape[0]
        kw = {}
        if hasattr(self.engine, "pinned_empty"):
            if self._bufs is None or self._bufs["C"] != C:
                from . import _lib
                pe = self.engine.pinned_empty
                self._bufs = {"C": C, "w": pe(weights.shape, np.float64), "res": pe((C * self.games,), _lib.RESULT_DTYPE),
                              "fit": pe((C,), np.float64), "pieces": None}
                if self.pieces is not None:
                    self._bufs["pieces"] = pe(self.pieces.shape, np.uint8)
                    np.copyto(self._bufs["pieces"], self.pieces)
            np.copyto(self._bufs["w"], weights)
            weights = self._bufs["w"]
            kw = {"out_results": self._bufs["res"], "out_fitness": self._bufs["fit"]}
        pieces = self._bufs["pieces"] if (self._bufs is not None and self._bufs["pieces"] is not None) else self.pieces
        out = self.engine.play_games(weights, self.games, pieces=pieces, T=self.T, seed=self.seed,
                                     lookahead=self.lookahead, lanes=self.lanes, **kw)
        self.last = out
        self.generations += 1
        return out["fitness"]


def shard_bounds(n, world, rank):
    """Contiguous slice [lo, hi) of n candidates owned by `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedEvaluator:
    """Candidate-sharded evaluation across the ranks of a torch.distributed process group (learning.py:38-41 on every rank).

    `local` is either
      * a PopulationEvaluator (GPU ranks): the rank's slice is played with a device-mode call on torch's current stream, the
        per-candidate fitness stays in HBM, ONE all_gather_into_tensor (NCCL over NVLink) collects f64[C] on every rank and ONE
        device->host copy brings it back; every buffer is allocated once and reused; or
      * a callable  weights f64[c][F] -> fitness f64[c]  (the CPU tests drive it with the oracle over gloo).
    Every rank must run the same seeded driver.  The all-gather carries one extra slot per rank with a checksum of the FULL
    weight batch, so a desynchronised driver raises instead of silently refitting different populations.
    """

    def __init__(self, local, group=None, device=None, names=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device
        self.pop = local if isinstance(local, PopulationEvaluator) else None
        self.local_eval = None if self.pop is not None else local
        self.names = list(self.pop.names) if self.pop is not None else ([getattr(f, "__name__", f) for f in names] if names is not None else None)
        self._batch = []
        self._keys_ok = None
        self._dev = None                      # device-path buffers, keyed by (C, F)
        self.last_bytes = {"h2d": 0, "d2h": 0}
        self.last_timing = {}                 # evaluate_ms / allgather_ms / d2h_ms (CUDA events), host_prep_ms, total_ms
        self.collect_timing = False           # when set, last_timing is filled from CUDA events around the play / the gather / the copy back
        self.collect_placements = False       # with collect_timing: last_timing["placements"] = scored leaves of the local shard (one small reduction)

    # -- the two callables handed to cross_entropy ---------------------------------------------
    def test_f(self, weights):
        keys = tuple(weights)
        if keys != self._keys_ok:
            if self.names is not None:
                got = [getattr(k, "__name__", k) for k in keys]
                if got != self.names:
                    raise ValueError("weights mapping keys %r do not match the evaluator's feature order %r" % (got, self.names))
            self._keys_ok = keys
        self._batch.append(list(weights.values()))
        return _Pending(len(self._batch) - 1)

    def map_f(self, fn, indices):
        self._batch = []
        tagged = [fn(i) for i in indices]
        fitness = self.evaluate(np.array(self._batch, dtype=np.float64)).tolist()
        return [(fitness[p.slot], i) for p, i in tagged]

    # -- batch evaluation ----------------------------------------------------------------------------
    @staticmethod
    def _checksum(weights):
        """48-bit digest of the whole batch as an exactly representable double."""
        import hashlib
        return float(int.from_bytes(hashlib.sha256(np.ascontiguousarray(weights, dtype=np.float64).tobytes()).digest()[:6], "little"))

    def _unpack(self, g, C):
        per = g.shape[1] - 1
        sums = g[:, per]
        if not (sums == sums[0]).all():
            raise RuntimeError("ShardedEvaluator: ranks evaluated different weight batches (checksums %r); every rank must run "
                               "the same seeded cross_entropy driver" % (sums.tolist(),))
        out = np.empty(C, dtype=np.float64)
        for r in range(self.world):
            a, b = shard_bounds(C, self.world, r)
            out[a:b] = g[r, : b - a]
        return out

    def evaluate(self, weights):
        weights = np.asarray(weights, dtype=np.float64)
        if self.pop is not None and self.device is not None and getattr(self.device, "type", "cpu") == "cuda":
            return self._evaluate_device(weights)
        return self._evaluate_host(weights)

    def _evaluate_host(self, weights):
        import torch
        C = weights.shape[0]
        lo, hi = shard_bounds(C, self.world, self.rank)
        ev = self.local_eval if self.local_eval is not None else self.pop.evaluate
        local = np.asarray(ev(weights[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
        per = -(-C // self.world)                                    # padded shard length
        buf = torch.zeros(per + 1, dtype=torch.float64, device=self.device)
        buf[: hi - lo] = torch.from_numpy(local).to(buf.device)
        buf[per] = self._checksum(weights)
        gathered = torch.empty((per + 1) * self.world, dtype=torch.float64, device=self.device)
        self.dist.all_gather_into_tensor(gathered, buf, group=self.group)
        return self._unpack(gathered.cpu().numpy().reshape(self.world, per + 1), C)

    def _buffers(self, C, F):
        import torch
        from . import _lib
        if self._dev is not None and self._dev["key"] == (C, F):
            return self._dev
        per = -(-C // self.world)
        G, dev, pop = self.pop.games, self.device, self.pop
        d = {"key": (C, F), "per": per,
             "h_w": torch.empty((per, F), dtype=torch.float64).pin_memory(),
             "d_w": torch.empty((per, F), dtype=torch.float64, device=dev),
             "d_res": torch.zeros(per * G * _lib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev),
             "d_fit": torch.zeros(per + 1, dtype=torch.float64, device=dev),
             "d_all": torch.empty((per + 1) * self.world, dtype=torch.float64, device=dev),
             "h_all": torch.empty((per + 1) * self.world, dtype=torch.float64).pin_memory(),
             "h_pieces": torch.from_numpy(pop.pieces).pin_memory() if pop.pieces is not None else None,
             "d_pieces": torch.empty(pop.pieces.shape, dtype=torch.uint8, device=dev) if pop.pieces is not None else None,
             "ev": [torch.cuda.Event(enable_timing=True) for _ in range(4)]}
        self._dev = d
        return d

    def _evaluate_device(self, weights):
        import time
        import torch
        t0 = time.perf_counter()
        pop, eng = self.pop, self.pop.engine
        C, F = weights.shape
        lo, hi = shard_bounds(C, self.world, self.rank)
        c = hi - lo
        b = self._buffers(C, F)
        per = b["per"]
        eng.set_features(pop.names)
        stream = torch.cuda.current_stream(self.device)
        timing = self.collect_timing
        ev = b["ev"]
        # ---- this step's inputs (the rank's weights, the injected piece sequences) and the play: enqueued first, so that the
        #      GPU is busy while the host does the rest of its bookkeeping ----
        b["h_w"].numpy()[:c] = weights[lo:hi]
        if timing:
            ev[0].record(stream)
        b["d_w"].copy_(b["h_w"], non_blocking=True)
        if b["d_pieces"] is not None:
            b["d_pieces"].copy_(b["h_pieces"], non_blocking=True)
        if c > 0:
            key = (stream.cuda_stream, c)
            if b.get("call_key") != key:                    # argument block of the device-mode call, built once per (stream, slice size)
                b["call"] = eng.prepare_device_call(stream.cuda_stream, c, pop.games, b["d_w"].data_ptr(), b["d_res"].data_ptr(), pop.T,
                                                    pieces_ptr=b["d_pieces"].data_ptr() if b["d_pieces"] is not None else 0,
                                                    n_sequences=pop.pieces.shape[0] if pop.pieces is not None else 0, seed=pop.seed,
                                                    lookahead=pop.lookahead, lanes=pop.lanes, fitness_ptr=b["d_fit"].data_ptr())
                b["call_key"] = key
            eng.run_prepared(b["call"])
        t1 = time.perf_counter()
        if timing:
            ev[1].record(stream)
        b["d_fit"][per:].fill_(self._checksum(weights))     # slot `per` of my block: checksum of the WHOLE batch (stream-ordered before the gather)
        self.dist.all_gather_into_tensor(b["d_all"], b["d_fit"], group=self.group)
        if timing:
            ev[2].record(stream)
        b["h_all"].copy_(b["d_all"], non_blocking=True)
        ev[3].record(stream)
        ev[3].synchronize()
        if c > 0:
            eng.check_status()
        out = self._unpack(b["h_all"].numpy().reshape(self.world, per + 1), C)
        self.last_bytes = {"h2d": c * F * 8 + (pop.pieces.size if pop.pieces is not None else 0), "d2h": (per + 1) * self.world * 8}
        if timing:
            self.last_timing = {"host_prep_ms": 1e3 * (t1 - t0), "evaluate_ms": ev[0].elapsed_time(ev[1]), "allgather_ms": ev[1].elapsed_time(ev[2]),
                                "d2h_ms": ev[2].elapsed_time(ev[3]), "total_ms": 1e3 * (time.perf_counter() - t0), "local_candidates": c}
            if self.collect_placements:
                from . import _lib
                rec = b["d_res"][: c * pop.games * _lib.RESULT_DTYPE.itemsize].view(torch.int64).view(-1, 5)
                self.last_timing["placements"] = int(rec[:, 4].sum().item())
        pop.generations += 1
        return out
