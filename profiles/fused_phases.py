"""CTA-level phase cycles of the LAST step of a fused launch (step_many, -DCYG_CTA_TIMING build) next to the launch time."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cygym_b200 import synthetic_network
from cygym_b200.vector_env import VectorCyberDefenseEnv
B = 65536
T = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net = synthetic_network(100, n_subnets=8, seed=0)
env = VectorCyberDefenseEnv(net, B, seed=0)
dbg = torch.zeros(B, dtype=torch.int64, device="cuda")
env.L.cyg_set_debug_cycles(env.h, C.c_void_p(dbg.data_ptr()))
names = ["init+issue TMA", "zero + barrier", "sort (+ wait load)", "phase A", "phase B", "phase C*", "last stores"]
def batches(T, last_mode):
    hs, ms = [], []
    for t in range(T):
        mode = (t + T - 1 + last_mode) & 1 if False else (t & 1)
        ab = env.sample_actions(mode)
        if mode == 0:
            ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
        hs.append(ab.hdr.clone()); ms.append(ab.mask.clone())
    return torch.stack(hs).contiguous(), torch.stack(ms).contiguous()
for rep in range(4):
    TT = T + (rep & 1)  # odd T ends on a defender turn, even T on an attacker turn
    h, m = batches(TT, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step_many(h, m); e1.record(); torch.cuda.synchronize()
    v = dbg[:147 * 8].view(147, 8).cpu().numpy().astype(np.int64)
    print(f"T={TT} launch {e0.elapsed_time(e1) * 1e3:.1f} us = {e0.elapsed_time(e1) * 1e3 / TT:.1f} us per step; last step ({'att' if (TT - 1) & 1 else 'def'}) median / max cycles per CTA:")
    d = np.diff(v[:, :7], axis=1)
    for i in range(1, 6):
        print(f"   {names[i]:20s} {int(np.median(d[:, i])):8d} {int(d[:, i].max()):8d}")
    print(f"   whole launch per CTA (t0 -> last mark): median {int(np.median(v[:, 6] - v[:, 0]))} max {int((v[:, 6] - v[:, 0]).max())} min {int((v[:, 6] - v[:, 0]).min())} cycles; per step {np.median(v[:, 6] - v[:, 0]) / TT:.0f}")
# are slow CTAs persistent?  per-CTA totals of consecutive launches
tot = []
for rep in range(4):
    h, m = batches(T, 0)
    env.step_many(h, m); torch.cuda.synchronize()
    v = dbg[:147 * 8].view(147, 8).cpu().numpy().astype(np.int64)
    tot.append((v[:, 6] - v[:, 0])[:146].astype(np.float64))
tot = np.array(tot)
print("corr of per-CTA launch totals between consecutive launches:", np.round(np.corrcoef(tot)[0], 3))
print("slowest CTAs per launch:", [list(np.argsort(-x)[:5]) for x in tot])
print("max/median per launch:", np.round(tot.max(axis=1) / np.median(tot, axis=1), 3))
smid = torch.zeros(1)
