"""BASELINE.json config C5: DOAR payoff-matrix evaluation, 32 x 32 strategy pairs x 1024 rollouts, sharded over the
ranks + ONE all-reduce of the [32, 32, 10] sums (NCCL).  Strategies: the reference's non-neural flavours (baseline
names and fixed action sequences, strategy.py:25-60).
    python profiles/c5_payoff.py                                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/c5_payoff.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cygym_b200 import synthetic_network  # noqa: E402
from cygym_b200.payoff import Strategy, evaluate_payoff_matrix_batched  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ND = NA = int(os.environ.get("C5_STRATEGIES", 32))
N, T = int(os.environ.get("C5_ROLLOUTS", 1024)), int(os.environ.get("C5_STEPS", 100))
net = synthetic_network(100, n_subnets=8, seed=0)
rng = np.random.default_rng(0)


def seq(mode, L):
    out = []
    for _ in range(L):
        n = int(rng.integers(1, 40))
        devs = sorted(int(d) for d in rng.choice(100, size=n, replace=False))
        at = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 11, 12, 13])) if mode == 0 else int(rng.integers(1, 3))
        out.append((at, [int(rng.integers(0, 2))], devs, int(rng.integers(0, 7))))
    return out


defs = [Strategy(baseline_name="No Defense"), Strategy(baseline_name="Preset"), Strategy(baseline_name="Nash")] + \
       [Strategy(actions=seq(0, 5)) for _ in range(ND - 3)]
atts = [Strategy(baseline_name="No Attack"), Strategy(baseline_name="Nash")] + [Strategy(actions=seq(1, 4)) for _ in range(NA - 2)]
# a tiny evaluation first: loads the CUDA module and sets the kernel attributes (about 1 s, paid once per process)
tw = time.perf_counter()
evaluate_payoff_matrix_batched(net, defs[:2], atts[:2], 8 * world, steps_per_episode=4, seed=1, device=dev, rank=rank, world=world)
torch.cuda.synchronize()
warm_s = time.perf_counter() - tw
t0 = time.perf_counter()
out = evaluate_payoff_matrix_batched(net, defs, atts, N, steps_per_episode=T, seed=1, device=dev, rank=rank, world=world)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"config": f"C5: {ND}x{NA} pairs x {N} rollouts x {T} steps", "n_gpus": world, "seconds": dt, "first_call_warmup_seconds": warm_s,
                      "env_steps_per_s": ND * NA * N * T / dt, "checksum": float(out.sum().item()),
                      "defender_return_mean": float(out[..., 0].mean().item()), "attacker_return_mean": float(out[..., 1].mean().item())}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
