import os, sys
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from cygym_b200 import synthetic_network
from cygym_b200.vector_env import VectorCyberDefenseEnv, ActionBatch
M, subnets, B, T = 100, 8, 5000, 7
net = synthetic_network(M, n_subnets=subnets, seed=5)
a = VectorCyberDefenseEnv(net, B, seed=11)
hdrs, masks = [], []
for t in range(T):
    ab = a.sample_actions(t & 1)
    if (t & 1) == 0:
        ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
    hdrs.append(ab.hdr.clone()); masks.append(ab.mask.clone()); a.step(ab)
hdr = torch.stack(hdrs).contiguous(); mask = torch.stack(masks).contiguous()
b = VectorCyberDefenseEnv(net, B, seed=11); d = VectorCyberDefenseEnv(net, B, seed=11)
raw, shaped, done = b.step_many(hdr, mask)
torch.cuda.synchronize()
for t in range(T):
    r = d.step(ActionBatch(hdr[t], mask[t]))
    torch.cuda.synchronize()
    bad = torch.nonzero(raw[t] != r[0]).flatten().cpu().numpy()
    at = (hdr[t][:, 0] & 0xFF).cpu().numpy()
    print(f"t={t} mode={t&1}: {len(bad)} envs differ; first {bad[:10]} types {at[bad[:10]]} env%64 {bad[:10] % 64}; type hist of bad {np.bincount(at[bad], minlength=14) if len(bad) else ''}")
print("NB", b.S)
print("---- test order: c single steps first, then b2.step_many")
b2 = VectorCyberDefenseEnv(net, B, seed=11); c = VectorCyberDefenseEnv(net, B, seed=11)
for t in range(T):
    c.step(ActionBatch(hdr[t], mask[t]))
raw2, shaped2, done2 = b2.step_many(hdr, mask)
torch.cuda.synchronize()
print("c vs d last raw:", torch.nonzero(c.raw != d.raw).flatten().cpu().numpy()[:12])
print("b2 vs d last raw:", torch.nonzero(raw2[T-1] != d.raw).flatten().cpu().numpy()[:12])
print("b2 vs b all rows:", [int((raw2[t] != raw[t]).sum()) for t in range(T)])
