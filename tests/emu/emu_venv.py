"""EmuVectorEnv: the call surface of cygym_b200.vector_env.VectorCyberDefenseEnv that the drop-in Gym class uses, on
top of the HOST compile of the device source (tests/emu).  TEST INFRASTRUCTURE: it lets the CPU-only build container --
the only place where /root/reference exists -- drive the drop-in class with the reference's unmodified callers
(tests/test_dropin_reference_caller.py).  The product has no CPU path; nothing under cygym_b200/ imports this."""
import numpy as np
import torch

from cygym_b200 import _capi as K
from cygym_b200.vector_env import ActionBatch
from oracle import cyg_oracle as O
from tests.emu import emu


class _Scalars:
    """`scalars[b, i] = v` / `scalars[b, i]` over the oracle-state array."""

    def __init__(self, st):
        self.st = st

    def __setitem__(self, idx, v):
        self.st.scal[idx] = np.uint32(int(v) & 0xFFFFFFFF)

    def __getitem__(self, idx):
        return torch.from_numpy(np.asarray(self.st.scal[idx]).astype(np.int64))


class EmuVectorEnv:
    def __init__(self, network, num_envs, device=None, seed=0, env_id0=0, base_line="Nash", xcap=16, stream=None, log_cap=0, detector_slots=0):
        self.net, self.B, self.M, self.W = network, int(num_envs), network.M, network.W
        self.xcap = max(int(xcap), len(network.template.get("extra", ())))
        self.cfg = O.make_config(network.cfg, network.E, seed=seed, xcap=self.xcap, base_line=base_line, log_cap=log_cap)
        self.emu = emu.Emu(dict(row_ptr=network.row_ptr, col=network.col, mult=network.mult, dev_static=network.dev_static,
                                os_val=network.os_val, ver_val=network.ver_val), self.cfg, env_id0=env_id0)
        self.st = O.OracleState(self.B, self.M, network.E, self.xcap, log_cap)
        self._orc = O.Oracle(dict(row_ptr=network.row_ptr, col=network.col, mult=network.mult, dev_static=network.dev_static,
                                  os_val=network.os_val, ver_val=network.ver_val), self.cfg, env_id0=env_id0)  # host-side detector service only
        self.scalars = _Scalars(self.st)
        self._out = torch.zeros(3, self.B, dtype=torch.float32)
        self._pre = torch.zeros(self.B, 3, self.W, dtype=torch.int32)
        self.reset()

    def close(self):
        pass

    def set_base_line(self, name):
        self.emu.set_base_line(name)

    def reset(self):
        t = self.net.template
        for k in ("dev", "ckpt", "blocked", "extra", "scal"):
            a = getattr(self.st, k)
            a[:] = 0
            src = np.asarray(t[k], np.uint32)
            a[:, : len(src)] = src[None]
        return self

    def import_state(self, canon):
        for k in ("dev", "ckpt", "blocked", "extra", "scal"):
            a = canon[k]
            a = a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)
            a = a.view(np.uint32).reshape(self.B, -1)
            dst = getattr(self.st, k)
            dst[:] = 0
            w = min(dst.shape[1], a.shape[1])
            dst[:, :w] = a[:, :w]
        if self.st.log_cap and canon.get("logs") is not None:
            a = canon["logs"]
            a = (a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)).view(np.uint32).reshape(self.B, -1)
            self.st.logs[:, : min(self.st.log_cap, a.shape[1])] = a[:, : self.st.log_cap]
        return self

    def export_state(self):
        d = {k: torch.from_numpy(getattr(self.st, k).view(np.int32).copy()) for k in ("dev", "ckpt", "blocked", "extra", "scal")}
        if self.st.log_cap:
            d["logs"] = torch.from_numpy(self.st.logs.view(np.int32).copy())
        return d

    def service_detectors(self, seed_of_env=None):
        return self._orc.service_detectors(self.st, seed_of_env)

    def to_device(self, hdr, mask, order=None):
        f = lambda a, dt: None if a is None else torch.from_numpy(np.ascontiguousarray(a).view(dt).copy())
        return ActionBatch(f(hdr, np.int32), f(mask, np.int32), f(order, np.int16))

    def _run(self, groups, flags):
        hdr = np.stack([g.hdr.numpy().view(np.uint32) for g in groups])
        mask = np.stack([g.mask.numpy().view(np.uint32) for g in groups])
        order = np.stack([g.order.numpy().view(np.uint16) for g in groups]) if groups[0].order is not None else None
        o = self.emu.step(self.st, hdr, mask, order, flags=flags, want_pre=True)
        self._out[0] = torch.from_numpy(o["raw"].astype(np.float32))
        self._out[1] = torch.from_numpy(o["shaped"].astype(np.float32))
        self._out[2] = torch.from_numpy(o["done"].astype(np.int32)).view(torch.float32)
        self._pre = torch.from_numpy(o["pre_masks"].view(np.int32).copy())
        return self._out[0], self._out[1], self._out[2].view(torch.int32)

    def step(self, actions, flags=0, obs_mode=0, want_pre=False):
        return self._run([actions], flags)

    def step_grouped(self, groups, obs_mode=0, want_pre=False):
        return self._run(list(groups), K.STEP_GROUPED)

    def pre_masks(self):
        return self._pre

    def observe(self, mode):
        return torch.from_numpy(self.emu.observe(self.st, mode))

    def randomize_compromise_and_ownership(self, env_mask=None):
        self.emu.randomize(self.st, env_mask)

    def rebuild_graph_cache(self, env_mask=None):
        self.emu.rebuild_graph_cache(self.st)

    def sample_actions(self, mode, out=None, want_order=False):
        m = 1 if mode in (1, "attacker") else 0
        h, mk, o = self.emu.sample_actions(self.st, m, want_order=True)
        ab = self.to_device(h, mk, o)
        return ab
