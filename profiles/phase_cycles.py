"""Where the cycles of one env's transition go (profiling build with -DCYG_PHASE_TIMING; diagnostics only).
Build:  nvcc <flags> -DCYG_PHASE_TIMING -o profiles/_build/libcygym_b200_prof.so cygym_b200/csrc/cyg_kernels.cu
Run  :  CYGYM_B200_LIB=profiles/_build/libcygym_b200_prof.so python profiles/phase_cycles.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cygym_b200 import synthetic_network  # noqa: E402
from cygym_b200.vector_env import VectorCyberDefenseEnv  # noqa: E402

B = 65536
net = synthetic_network(100, n_subnets=8, seed=0)
env = VectorCyberDefenseEnv(net, B, seed=0)
if os.environ.get("RANDOMIZE"):
    env.randomize_compromise_and_ownership()
dbg = torch.zeros(B * 8, dtype=torch.int64, device="cuda")
env.L.cyg_set_debug_cycles(env.h, C.c_void_p(dbg.data_ptr()))
names = ["epoch", "decode+tick", "action", "work", "arrivals", "count+reward", "evolve", "end"]
for t in range(40):
    mode = t & 1
    ab = env.sample_actions(mode)
    if mode == 0:
        ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
    env.step(ab)
    torch.cuda.synchronize()
    if t >= 36:
        ph = dbg.view(B, 8).cpu().numpy().astype(np.int64)
        at = (ab.hdr[:, 0] & 0xFF).cpu().numpy()
        d = np.diff(np.concatenate([np.zeros((B, 1), np.int64), ph], axis=1), axis=1)
        print(f"t={t} mode={'att' if mode else 'def'}: median cycles per phase " + " | ".join(names))
        for a in sorted(set(at.tolist())):
            m = at == a
            print(f"   type {a:2d}: " + " ".join(f"{int(np.median(d[m, i])):7d}" for i in range(8)) + f"   total {int(np.median(ph[m, 7]))}")
