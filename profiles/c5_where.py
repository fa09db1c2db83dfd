"""Where the time of one C5 evaluation goes (1 GPU): env construction, reset, randomize, table upload, rollout, reduction."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cygym_b200 import synthetic_network
from cygym_b200 import payoff as P
from cygym_b200.vector_env import VectorCyberDefenseEnv
net = synthetic_network(100, n_subnets=8, seed=0)
dev = "cuda:0"
def t(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); print(f"{label:28s} {1e3 * (time.perf_counter() - t0):8.2f} ms"); return r
for rep in range(2):
    print("rep", rep)
    n = 32 * 32 * 1024
    env = t("VectorCyberDefenseEnv(1M)", lambda: VectorCyberDefenseEnv(net, n, device=dev, seed=1, env_id0=0, xcap=16))
    t("reset()", env.reset)
    t("randomize", env.randomize_compromise_and_ownership)
    hdr = torch.zeros(100, 1024, 4, dtype=torch.int32, device=dev); hdr[:, :, 0] = 8 | (1 << 16); hdr[1::2, :, 0] = 3 | (1 << 8) | (1 << 16)
    mask = torch.zeros(100, 1024, 4, dtype=torch.int32, device=dev)
    bl = torch.zeros(100, 1024, dtype=torch.uint8, device=dev)
    ret = t("rollout (no-op tables)", lambda: env.rollout(hdr, mask, bl, 1024, 0))
    t("info + index_add", lambda: torch.zeros(1024, 10, dtype=torch.float64, device=dev).index_add_(0, torch.arange(n, device=dev) // 1024, torch.stack([ret[0], ret[1]] + [v.double() for v in list(env.info().values())[:8]], 1)))
    t("close", env.close)
    del env
