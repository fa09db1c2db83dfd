"""Shared helpers for the parity tests (TEST INFRASTRUCTURE)."""
import glob
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
GOLDEN_IDS = [os.path.basename(f)[:-4] for f in GOLDEN]


def load_golden(path):
    return dict(np.load(path))


class CudaImpl:
    """oracle.trajectory.replay() protocol on top of the CUDA library (B = 1), through the C-ABI."""

    def __init__(self, g):
        import torch
        from cygym_b200 import network_from_golden
        from cygym_b200.vector_env import VectorCyberDefenseEnv
        self.torch = torch
        net, meta = network_from_golden(g)
        self.meta = meta
        self.env = VectorCyberDefenseEnv(net, 1, seed=meta["draw_seed"], env_id0=meta["env_id"], xcap=meta["xcap"],
                                         log_cap=meta.get("log_cap", 0), detector_slots=1 if meta.get("log_cap", 0) else 0)

    def load(self, init):
        self.env.import_state({k: np.asarray(v, np.uint32)[None] for k, v in init.items()})

    def bump_epoch(self, n):
        if n:
            self.env.scalars[:, 1] += int(n)

    def sample_action(self, mode):
        ab = self.env.sample_actions(mode, want_order=True)
        self.torch.cuda.synchronize()
        return (ab.hdr.cpu().numpy().view(np.uint32)[0], ab.mask.cpu().numpy().view(np.uint32)[0],
                ab.order.cpu().numpy().view(np.uint16)[0])

    def step(self, hdr, mask, order, flags):
        env = self.env
        G = hdr.shape[0]
        from cygym_b200.vector_env import ActionBatch
        groups = []
        for g in range(G):
            groups.append(env.to_device(hdr[g], mask[g], None if order is None else order[g]))
        if flags & 1:
            raw, shaped, done = env.step_grouped(groups, want_pre=True)
        else:
            raw, shaped, done = env.step(groups[0], flags=flags, want_pre=True)
        self.torch.cuda.synchronize()
        return dict(raw=raw.cpu().numpy().astype(np.float64), shaped=shaped.cpu().numpy().astype(np.float64),
                    done=done.cpu().numpy(), pre_masks=env.pre_masks().cpu().numpy().view(np.uint32))

    def randomize(self):
        self.env.randomize_compromise_and_ownership()

    def set_base_line(self, name):
        self.env.set_base_line(name)

    def state(self):
        c = self.env.export_state()
        self.torch.cuda.synchronize()
        d = {k: v.cpu().numpy().view(np.uint32)[0] for k, v in c.items()}
        if "logs" in d:
            from oracle.trajectory import log_tail_of_ring
            d["logs_tail"] = log_tail_of_ring(d["logs"], int(d["scal"][6]), self.env.log_cap)
        return d

    def service_detector(self, seed):
        self.env.service_detectors(lambda b: seed)

    def observe(self, mode):
        o = self.env.observe(mode)
        self.torch.cuda.synchronize()
        return o.cpu().numpy()[0]


def oracle_for(net, seed, xcap, env_id0=0, base_line="Nash"):
    from oracle import cyg_oracle as O
    cfg = O.make_config(net.cfg, net.E, seed=seed, xcap=xcap, base_line=base_line)
    d = dict(row_ptr=net.row_ptr, col=net.col, mult=net.mult, dev_static=net.dev_static, os_val=net.os_val, ver_val=net.ver_val)
    return O.Oracle(d, cfg, env_id0=env_id0), cfg


def oracle_state_from_template(orc, net, B):
    st = orc.new_state(B)
    t = net.template
    st.dev[:] = np.asarray(t["dev"], np.uint32)[None]
    st.ckpt[:] = np.asarray(t["ckpt"], np.uint32)[None]
    st.blocked[:] = 0
    st.blocked[:, :len(t["blocked"])] = np.asarray(t["blocked"], np.uint32)[None]
    st.extra[:] = 0
    if len(t["extra"]):
        st.extra[:, :len(t["extra"])] = np.asarray(t["extra"], np.uint32)[None]
    st.scal[:] = np.asarray(t["scal"], np.uint32)[None]
    return st


def sanitize_actions(hdr, scal_logs, mode):
    """Drop the trained-IsolationForest branch (defender action 10 with a non-empty log): out of the
    kernel's scope.  Replaced by the no-op 8, in place."""
    if mode == 0:
        at = hdr[:, 0] & 0xFF
        bad = (at == 10) & (scal_logs > 0)
        hdr[bad, 0] = (hdr[bad, 0] & ~np.uint32(0xFF)) | np.uint32(8)
    return hdr


def compare_states(a, b, label, rtol=1e-5):
    """a, b: dicts of canonical numpy arrays [B, *]; raises AssertionError on the first mismatch."""
    for k in ("dev", "ckpt", "blocked", "extra"):
        x, y = np.asarray(a[k]).view(np.uint32), np.asarray(b[k]).view(np.uint32)
        w = min(x.shape[1], y.shape[1])
        if not np.array_equal(x[:, :w], y[:, :w]):
            bad = np.argwhere(x[:, :w] != y[:, :w])[:5]
            raise AssertionError(f"{label}: {k} differs at (env, idx) {bad.tolist()}: "
                                 f"{[hex(int(x[i, j])) for i, j in bad]} vs {[hex(int(y[i, j])) for i, j in bad]}")
    sa, sb = np.array(a["scal"]).view(np.uint32).copy(), np.array(b["scal"]).view(np.uint32).copy()
    fa, fb = sa[:, 9:11].copy().view(np.float32), sb[:, 9:11].copy().view(np.float32)
    if not np.allclose(fa, fb, rtol=rtol, atol=1e-6):
        raise AssertionError(f"{label}: cost counters differ")
    sa[:, 9:11] = 0
    sb[:, 9:11] = 0
    if not np.array_equal(sa, sb):
        bad = np.argwhere(sa != sb)[:5]
        raise AssertionError(f"{label}: scalars differ at (env, slot) {bad.tolist()}: "
                             f"{[int(sa[i, j]) for i, j in bad]} vs {[int(sb[i, j]) for i, j in bad]}")


def compare_rewards(out_a, out_b, label, rtol=1e-5):
    for k in ("raw", "shaped"):
        x, y = np.asarray(out_a[k], np.float64), np.asarray(out_b[k], np.float64)
        tol = rtol * np.maximum(1.0, np.abs(y))
        if np.any(np.abs(x - y) > tol):
            i = int(np.argmax(np.abs(x - y) - tol))
            raise AssertionError(f"{label}: {k} reward env {i}: {x[i]} vs {y[i]}")
    if not np.array_equal(np.asarray(out_a["done"]), np.asarray(out_b["done"])):
        raise AssertionError(f"{label}: done flags differ")
