import os, sys
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from cygym_b200 import synthetic_network
from cygym_b200.vector_env import VectorCyberDefenseEnv
for M, subnets, B, T in ((100, 8, 5000, 7), (30, 2, 777, 5)):
    net = synthetic_network(M, n_subnets=subnets, seed=5)
    a = VectorCyberDefenseEnv(net, B, seed=11)
    hdrs, masks = [], []
    for t in range(T):
        ab = a.sample_actions(t & 1)
        if (t & 1) == 0:
            ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
        hdrs.append(ab.hdr.clone()); masks.append(ab.mask.clone())
        a.step(ab)
    b = VectorCyberDefenseEnv(net, B, seed=11)
    c = VectorCyberDefenseEnv(net, B, seed=11)
    hdr = torch.stack(hdrs).contiguous(); mask = torch.stack(masks).contiguous()
    rows = []
    for t in range(T):
        r = c.step(type(ab)(hdr[t], mask[t]))
        rows.append(r[0].clone())
    raw, shaped, done = b.step_many(hdr, mask)
    torch.cuda.synchronize()
    for t in range(T):
        bad = torch.nonzero(raw[t] != rows[t]).flatten().cpu().numpy()
        at = (hdr[t][:, 0] & 0xFF).cpu().numpy()
        print(M, f"t={t}: {len(bad)} differ; first {bad[:10]} types {at[bad[:10]]}; hist {np.bincount(at[bad], minlength=14) if len(bad) else ''}")
