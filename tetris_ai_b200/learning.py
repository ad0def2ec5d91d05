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
    """Candidate-sharded evaluation across the ranks of a torch.distributed process group.

    local_eval(weights f64[c][F]) -> fitness f64[c] is the per-rank evaluator (on a GPU box:
    PopulationEvaluator.evaluate).  The only collective is an all-gather of the fitness vector
    (NCCL over NVLink when the group's backend is nccl; gloo in the CPU tests).
    """

    def __init__(self, local_eval, group=None, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.local_eval = local_eval
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device
        self._batch = []
        self.names = None

    def test_f(self, weights):
        self._batch.append(list(weights.values()))
        return _Pending(len(self._batch) - 1)

    def map_f(self, fn, indices):
        self._batch = []
        tagged = [fn(i) for i in indices]
        fitness = self.evaluate(np.array(self._batch, dtype=np.float64)).tolist()
        return [(fitness[p.slot], i) for p, i in tagged]

    def evaluate(self, weights):
        import torch
        C = weights.shape[0]
        lo, hi = shard_bounds(C, self.world, self.rank)
        local = np.asarray(self.local_eval(weights[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
        per = -(-C // self.world)                                    # padded shard length
        buf = torch.zeros(per, dtype=torch.float64, device=self.device)
        buf[: hi - lo] = torch.from_numpy(local).to(buf.device)
        gathered = torch.empty(per * self.world, dtype=torch.float64, device=self.device)
        self.dist.all_gather_into_tensor(gathered, buf, group=self.group)
        g = gathered.cpu().numpy().reshape(self.world, per)
        out = np.empty(C, dtype=np.float64)
        for r in range(self.world):
            a, b = shard_bounds(C, self.world, r)
            out[a:b] = g[r, : b - a]
        return out
