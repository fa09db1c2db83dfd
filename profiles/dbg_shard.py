import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from cygym_b200 import synthetic_network
from cygym_b200.vector_env import VectorCyberDefenseEnv
net = synthetic_network(100, n_subnets=8, seed=0)
B, T, xcap = 65536, 24, 16
whole = VectorCyberDefenseEnv(net, B, seed=42, xcap=xcap)
halves = [VectorCyberDefenseEnv(net, B // 2, seed=42, env_id0=i * (B // 2), xcap=xcap) for i in range(2)]
def st(e): return {k: v.cpu().numpy().view(np.uint32) for k, v in e.export_state().items()}
for t in range(T):
    mode = t & 1
    ab = whole.sample_actions(mode)
    hs = [h.sample_actions(mode) for h in halves]
    torch.cuda.synchronize()
    if mode == 0:
        for a, e in [(ab, whole)] + list(zip(hs, halves)):
            bad = ((a.hdr[:, 0] & 0xFF) == 10) & (e.scalars[:, 6] > 0)
            a.hdr[:, 0] = torch.where(bad, (a.hdr[:, 0] & ~0xFF) | 8, a.hdr[:, 0])
    whole.step(ab)
    for h, a in zip(halves, hs): h.step(a)
    torch.cuda.synchronize()
    cw = st(whole); ch = [st(h) for h in halves]
    for k in cw:
        cat = np.concatenate([ch[0][k], ch[1][k]])
        if not np.array_equal(cat, cw[k]):
            d = np.where((cat != cw[k]).reshape(B, -1).any(axis=1))[0]
            at = (ab.hdr[:, 0] & 0xFF).cpu().numpy(); nd = ab.hdr[:, 2].cpu().numpy()
            print(f"t={t} mode={mode} key={k}: {len(d)} envs differ; first {d[:12]}; types {at[d[:12]]}; n_dev {nd[d[:12]]}; env%448 {d[:12]%448}; env%224 {d[:12]%224}")
            for e in d[:3]:
                print("   whole", cw[k][e][:18], "\n   halves", cat[e][:18], "\n   xor", (cw[k][e] ^ cat[e])[:18])
    if t >= 23: break
print("done")
