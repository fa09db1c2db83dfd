"""ctypes binding of tests/emu/cyg_emu.cpp: the DEVICE source compiled for the host (TEST INFRASTRUCTURE).

Lets the CPU-only container replay golden trajectories through the very code the CUDA kernels
run.  Not a product path: cygym_b200 never imports this.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_SO = os.path.join(_HERE, "_build", "libcyg_emu.so")
_SRCS = [os.path.join(_HERE, "cyg_emu.cpp"), os.path.join(_ROOT, "cygym_b200", "csrc", "cyg_core.cuh"),
         os.path.join(_ROOT, "cygym_b200", "csrc", "cyg_tables.h"), os.path.join(_ROOT, "include", "cygym_b200.h")]


def build(force=False):
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in _SRCS):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", _SO, _SRCS[0]],
                   check=True, capture_output=True, text=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.c_void_p] * 7
        L.emu_last_error.restype = C.c_char_p
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_set_base_line.argtypes = [C.c_void_p, C.c_int32]
        L.emu_record_words.argtypes = [C.c_void_p]
        L.emu_step.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_uint32] + [C.c_void_p] * 5
        L.emu_randomize.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.emu_set_aux.argtypes = [C.c_void_p] * 4
        L.emu_pyset_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.emu_rebuild.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.emu_sample_actions.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.emu_observe.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Emu:
    """Same call surface as oracle.cyg_oracle.Oracle, on oracle.cyg_oracle.OracleState arrays."""

    def __init__(self, netw, config, env_id0=0):
        self.L = lib()
        self.cfg = config
        self.M, self.E, self.X = config.M, config.E, config.X
        self.W = (self.M + 31) // 32
        self.env_id0 = env_id0
        self._keep = [np.ascontiguousarray(netw["row_ptr"], np.int32), np.ascontiguousarray(netw["col"], np.int32),
                      np.ascontiguousarray(netw["mult"], np.uint8), np.ascontiguousarray(netw["dev_static"], np.uint32),
                      np.ascontiguousarray(netw["os_val"], np.float32), np.ascontiguousarray(netw["ver_val"], np.float32)]
        self.h = self.L.emu_create(C.addressof(config), *[_p(a) for a in self._keep])
        if not self.h:
            raise ValueError(self.L.emu_last_error().decode())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.emu_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_base_line(self, name):
        from oracle.cyg_oracle import BL
        self.L.emu_set_base_line(self.h, BL.get(name, 4))

    def step(self, st, hdr, mask, order=None, flags=0, want_pre=False, **_):
        hdr = np.ascontiguousarray(hdr, np.uint32)
        mask = np.ascontiguousarray(mask, np.uint32)
        if hdr.ndim == 2:
            hdr, mask = hdr[None], mask[None]
            if order is not None:
                order = order[None]
        G, B = hdr.shape[0], hdr.shape[1]
        ostride = 0
        if order is not None:
            order = np.ascontiguousarray(order, np.uint16)
            ostride = order.shape[2]
        raw = np.zeros(B, np.float32); shaped = np.zeros(B, np.float32)
        done = np.zeros(B, np.int32); ex = np.zeros(B, np.int32)
        pre = np.zeros((B, 3, self.W), np.uint32) if want_pre else None
        self._keep_aux = (st.logs, st.det_slots, st.det_of_env)
        self.L.emu_set_aux(self.h, _p(st.logs) if getattr(st, "log_cap", 0) > 0 else None,
                           _p(st.det_slots) if st.det_slots is not None else None, _p(st.det_of_env))
        self.L.emu_step(self.h, B, self.env_id0, _p(st.dev), _p(st.ckpt), _p(st.blocked), _p(st.extra), _p(st.scal),
                        _p(hdr), _p(mask), _p(order), ostride, G, flags, _p(raw), _p(shaped), _p(done), _p(ex), _p(pre))
        out = dict(raw=raw, shaped=shaped, done=done, exec_atype=ex)
        if want_pre:
            out["pre_masks"] = pre
        return out

    def randomize(self, st, env_mask=None):
        if env_mask is not None:
            env_mask = np.ascontiguousarray(env_mask, np.uint8)
        self.L.emu_randomize(self.h, st.B, self.env_id0, _p(st.dev), _p(st.blocked), _p(st.extra), _p(st.scal), _p(env_mask))

    def rebuild_graph_cache(self, st):
        self.L.emu_rebuild(self.h, st.B, self.env_id0, _p(st.dev), _p(st.blocked), _p(st.extra), _p(st.scal))

    def sample_actions(self, st, mode, want_order=False):
        hdr = np.zeros((st.B, 4), np.uint32)
        mask = np.zeros((st.B, self.W), np.uint32)
        order = np.zeros((st.B, self.M), np.uint16) if want_order else None
        self.L.emu_sample_actions(self.h, st.B, self.env_id0, _p(st.scal), mode, _p(hdr), _p(mask), _p(order), self.M)
        return (hdr, mask, order) if want_order else (hdr, mask)

    def observe(self, st, obs_mode):
        dim = 4 * self.M + self.X if obs_mode == 2 else 6 * self.M
        obs = np.zeros((st.B, dim), np.float32)
        self.L.emu_observe(self.h, st.B, _p(st.dev), obs_mode, _p(obs))
        return obs


class EmuImpl:
    """oracle.trajectory.replay() protocol on top of the host build of the device source."""

    def __init__(self, g):
        from oracle import cyg_oracle as O
        meta = json.loads(str(g["meta"]))
        netw = {k: np.array(g["net_" + k]) for k in ("row_ptr", "col", "mult", "dev_static", "os_val", "ver_val")}
        self.cfg = O.make_config(meta["cfg"], len(netw["col"]), seed=meta["draw_seed"], xcap=meta["xcap"], log_cap=meta.get("log_cap", 0))
        self.emu = Emu(netw, self.cfg, env_id0=meta["env_id"])
        self.st = O.OracleState(1, self.cfg.M, self.cfg.E, self.cfg.xcap, self.cfg.log_cap)
        self._orc = O.Oracle(netw, self.cfg, env_id0=meta["env_id"])  # only its host-side detector service (sklearn fit + pack)

    def load(self, init):
        self.st.dev[0] = init["dev"]; self.st.ckpt[0] = init["ckpt"]
        self.st.blocked[0] = 0; self.st.blocked[0, :len(init["blocked"])] = init["blocked"]
        self.st.extra[0] = 0; self.st.extra[0, :len(init["extra"])] = init["extra"]
        self.st.scal[0] = init["scal"]

    def bump_epoch(self, n):
        self.st.scal[0, 1] += np.uint32(n)

    def sample_action(self, mode):
        h, m, o = self.emu.sample_actions(self.st, mode, want_order=True)
        return h[0], m[0], o[0]

    def step(self, hdr, mask, order, flags):
        return self.emu.step(self.st, hdr, mask, order, flags=flags, want_pre=True)

    def randomize(self):
        self.emu.randomize(self.st)

    def set_base_line(self, name):
        self.emu.set_base_line(name)

    def state(self):
        d = dict(dev=self.st.dev[0], ckpt=self.st.ckpt[0], blocked=self.st.blocked[0], extra=self.st.extra[0], scal=self.st.scal[0])
        if self.st.log_cap > 0:
            from oracle.trajectory import log_tail_of_ring
            d["logs_tail"] = log_tail_of_ring(self.st.logs[0], int(self.st.scal[0, 6]), self.st.log_cap)
        return d

    def service_detector(self, seed):
        self._orc.service_detectors(self.st, lambda b: seed)

    def observe(self, mode):
        return self.emu.observe(self.st, mode)[0]
