"""CTA-level phase times of the step kernel (profiling build -DCYG_CTA_TIMING; diagnostics only).
Run on the GPU box:  nvcc ... -DCYG_CTA_TIMING -o /tmp/cta.so cygym_b200/csrc/cyg_kernels.cu &&
                     CYGYM_B200_LIB=/tmp/cta.so python profiles/cta_phases.py"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cygym_b200 import synthetic_network
from cygym_b200.vector_env import VectorCyberDefenseEnv
B = 65536
net = synthetic_network(100, n_subnets=8, seed=0)
env = VectorCyberDefenseEnv(net, B, seed=0)
dbg = torch.zeros(B, dtype=torch.int64, device="cuda")
env.L.cyg_set_debug_cycles(env.h, C.c_void_p(dbg.data_ptr()))
names = ["init+issue TMA", "sort + wait load", "phase A", "phase B", "phase C", "obs/fence", "store"]
for t in range(40):
    mode = t & 1
    ab = env.sample_actions(mode)
    if mode == 0:
        ab.hdr[:, 0] = torch.where((ab.hdr[:, 0] & 0xFF) == 10, (ab.hdr[:, 0] & ~0xFF) | 8, ab.hdr[:, 0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step(ab); e1.record(); torch.cuda.synchronize()
    if t >= 38:
        v = dbg[:147 * 8].view(147, 8).cpu().numpy().astype(np.int64)
        d = np.diff(v[:, :7 + 1][:, :7], axis=1) if False else np.diff(v[:, :7], axis=1)
        print(f"t={t} {'att' if mode else 'def'} launch {e0.elapsed_time(e1) * 1e3:.1f} us; median / max cycles per CTA phase:")
        for i in range(6):
            print(f"   {names[i]:18s} {int(np.median(d[:, i])):8d} {int(d[:, i].max()):8d}")
        wm = dbg[2048:2048 + 147 * 32 * 4].view(147, 32, 4).cpu().numpy().astype(np.int64)[:, :28, :3]
        t3 = v[:, 3][:, None]  # phase B start
        for i, nm in enumerate(["B1 block/unblock", "B2 deposits", "B3 attack"]):
            endw = wm[:, :, i] - t3
            print(f"   per-warp end of {nm:18s} (cycles after phase-B start): median {int(np.median(endw)):8d}  min {int(endw.min()):8d}  CTA-max median {int(np.median(endw.max(axis=1))):8d}")
        print(f"   total              {int(np.median(v[:, 6] - v[:, 0])):8d} {int((v[:, 6] - v[:, 0]).max()):8d}  ({np.median(v[:,6]-v[:,0])/1.965e3:.1f} us)")
