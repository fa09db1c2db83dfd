"""Per-action-type cost of the step kernel: SM cycles each env's transition took, grouped by the action type
it executed (diagnostics; uses cyg_set_debug_cycles).  Run on the GPU box: python profiles/type_cycles.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cygym_b200 import synthetic_network  # noqa: E402
from cygym_b200.vector_env import VectorCyberDefenseEnv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
net = synthetic_network(100, n_subnets=8, seed=0)
env = VectorCyberDefenseEnv(net, B, seed=0)
if os.environ.get("RANDOMIZE"):  # extra hub-star edges in every env (the Double-Oracle rollout case)
    env.randomize_compromise_and_ownership()
dbg = torch.zeros(B, dtype=torch.int64, device="cuda")
env.L.cyg_set_debug_cycles(env.h, C.c_void_p(dbg.data_ptr()))
for t in range(40):
    mode = t & 1
    ab = env.sample_actions(mode)
    if mode == 0:
        ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step(ab)
    e1.record()
    torch.cuda.synchronize()
    if t >= 36:
        cyc = dbg.cpu().numpy()
        at = (ab.hdr[:, 0] & 0xFF).cpu().numpy()
        nd = (ab.hdr[:, 2] & 0xFFFF).cpu().numpy()
        print(f"t={t} mode={'att' if mode else 'def'} launch {e0.elapsed_time(e1) * 1e3:.1f} us; per type: mean / p50 / max cycles (n, mean n_dev)")
        for a in sorted(set(at.tolist())):
            m = at == a
            c = cyc[m]
            print(f"   type {a:2d}: {c.mean():9.0f} {np.median(c):9.0f} {c.max():9.0f}   ({m.sum()}, {nd[m].mean():.0f})")
        print(f"   all    : {cyc.mean():9.0f} {np.median(cyc):9.0f} {cyc.max():9.0f}   sum/SM = {cyc.sum() / 148 / 32:.0f} warp-cycles if perfectly packed")
