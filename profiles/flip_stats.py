"""Block / unblock task statistics (diagnostics).  With the default library: SM cycles per task by listed-device count.
With a -DCYG_COUNT_ROUNDS build (CYGYM_B200_LIB=...): windows (low 16 bits) and fixed-point passes (high bits) per task."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from cygym_b200 import synthetic_network
from cygym_b200.vector_env import VectorCyberDefenseEnv
B = 65536
rounds = "rounds" in os.environ.get("CYGYM_B200_LIB", "")
net = synthetic_network(100, n_subnets=8, seed=0)
env = VectorCyberDefenseEnv(net, B, seed=0)
dbg = torch.zeros(B, dtype=torch.int64, device='cuda')
env.L.cyg_set_debug_cycles(env.h, C.c_void_p(dbg.data_ptr()))
for t in range(39):
    mode = t & 1
    ab = env.sample_actions(mode)
    if mode == 0:
        ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
    env.step(ab); torch.cuda.synchronize()
v = dbg.cpu().numpy(); at = (ab.hdr[:, 0] & 0xFF).cpu().numpy(); nd = ab.hdr[:, 2].cpu().numpy()
for a in (6, 9):
    m = at == a; x = v[m]; n = nd[m]
    if rounds:
        wins, passes = x & 0xFFFF, x >> 16
        print('type', a, 'windows pcts', [int(np.percentile(wins, p)) for p in (10, 50, 90, 99, 100)],
              'passes pcts', [int(np.percentile(passes, p)) for p in (10, 50, 75, 90, 95, 99, 100)], 'mean passes', passes.mean())
        x = passes
    else:
        print('type', a, 'cycles pcts', [int(np.percentile(x, p)) for p in (10, 50, 75, 90, 95, 99, 100)])
    for lo, hi in ((1, 16), (16, 32), (32, 48), (48, 64), (64, 80), (80, 91)):
        mm = (n >= lo) & (n < hi)
        if mm.any():
            print('   n_dev', lo, hi, 'mean', int(x[mm].mean()), 'p90', int(np.percentile(x[mm], 90)), 'max', int(x[mm].max()))
