"""Multi-GPU check (run under torchrun on a box with >= 2 GPUs, see tests/test_gpu_multi.py):
sharded CE generations over NCCL must equal the single-process golden run bit for bit on every rank."""
import json
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_ai_b200 import Engine                                         # noqa: E402
from tetris_ai_b200.learning import PopulationEvaluator, ShardedEvaluator, cross_entropy   # noqa: E402
from tetris_ai_b200.tetris import features as F                           # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ce = json.load(open(os.path.join(ROOT, "tests", "golden", "ce.json")))["ce"]
    feats = [getattr(F, n) for n in ce["features"]]
    seqs = []
    for g in range(ce["games"]):
        r = random.Random(ce["seed0"] + g)
        seqs.append([r.randrange(7) for _ in range(ce["T"])])
    eng = Engine(local)
    local_ev = PopulationEvaluator(eng, feats, ce["games"], pieces=np.array(seqs, dtype=np.uint8))
    ev = ShardedEvaluator(local_ev, device=dev)          # device path: play -> all-gather of the device fitness -> one D2H
    gen = cross_entropy(feats, ce["stdev"], ce["width"], ce["keep"], random.Random(ce["rng_seed"]), ev.test_f, map_f=ev.map_f)
    ok = True
    for want in ce["generations"]:
        mean, std = next(gen)
        ok &= [float(v).hex() for v in mean.values()] == want["mean"]
        ok &= [float(v).hex() for v in std.values()] == want["std"]
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if dist.get_rank() == 0:
        print("DIST_CE_OK" if int(flag.item()) == 1 else "DIST_CE_MISMATCH", "world", dist.get_world_size())
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
