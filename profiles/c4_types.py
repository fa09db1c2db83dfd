"""C4 (1024 envs x 2000 device slots): launch time per turn and per action type (all envs forced to one type)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cygym_b200 import synthetic_network
from cygym_b200.vector_env import VectorCyberDefenseEnv
B = 1024
net = synthetic_network(2000, n_subnets=64, seed=0)
env = VectorCyberDefenseEnv(net, B, seed=0)
def t_launch(ab, n=3):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.step(ab); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)
for mode in (0, 1):
    ab = env.sample_actions(mode)
    ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
    print("mode", mode, "mixed types: %.0f us" % t_launch(ab))
    for ty in range(14 if mode == 0 else 5):
        if mode == 0 and ty == 10: continue
        h = ab.hdr.clone(); h[:, 0] = (h[:, 0] & ~0xFF) | ty
        from cygym_b200.vector_env import ActionBatch
        print("   type %2d: %8.0f us" % (ty, t_launch(ActionBatch(h, ab.mask))))
