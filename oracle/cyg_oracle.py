"""ctypes binding of the CPU checker oracle/cyg_oracle.c (TEST INFRASTRUCTURE ONLY).

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs only.  Never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import draws as D

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcyg_oracle.so")

NSCAL = 16
ATYPE_NONE = 0x80
STEP_GROUPED, STEP_SKIP_WORK = 1, 2
BL = {"Nash": 0, "No Defense": 1, "Preset": 2, "No Attack": 3}


class CygConfig(C.Structure):
    """Mirror of `struct cyg_config` in include/cygym_b200.h."""
    _fields_ = [
        ("M", C.c_int32), ("E", C.c_int32), ("X", C.c_int32), ("n_exploits", C.c_int32), ("xcap", C.c_int32),
        ("num_of_device", C.c_int32), ("min_network_size", C.c_int32), ("evolve_period", C.c_int32),
        ("wl_period_base", C.c_int32), ("wl_period_max", C.c_int32), ("wl_cap", C.c_int32),
        ("scaling_vulnerability", C.c_int32), ("turbo", C.c_int32), ("zero_day", C.c_int32),
        ("zero_day_mask", C.c_uint32), ("att_space_n", C.c_int32), ("def_space_n", C.c_int32),
        ("default_high", C.c_int32), ("n_app_ids", C.c_int32), ("base_line", C.c_int32), ("tri_high", C.c_int32),
        ("log_cap", C.c_int32),
        ("work_scale", C.c_float), ("comp_scale", C.c_float), ("def_scale", C.c_float), ("gamma", C.c_float),
        ("thr_p_add", C.c_uint64), ("thr_p_attacker", C.c_uint64),
        ("poisson_tab", C.c_uint32 * 16), ("tri_tab", C.c_uint32 * 8), ("seed", C.c_uint64),
        ("turbo_frac_clients", C.c_double), ("turbo_frac_servers", C.c_double), ("turbo_max_clients", C.c_int32),
        ("turbo_max_servers", C.c_int32), ("turbo_ramp_steps", C.c_int32), ("reserved1", C.c_int32),
    ]


def build():
    """Compile the checker (make -C oracle).  Building the checker is not using it."""
    try:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True, text=True)
    except subprocess.CalledProcessError:
        subprocess.run(["make", "-C", _HERE, "OMP="], check=True, capture_output=True, text=True)
    return _SO


def make_config(cfg, E, seed=0, xcap=16, base_line="Nash", tri_mode=2, tri_high=5, n_app_ids=0, log_cap=0):
    """cfg: the dict produced by ref_harness.extract_network()/the network generator."""
    c = CygConfig()
    c.M, c.E, c.X, c.n_exploits, c.xcap = cfg["M"], E, cfg["X"], cfg["n_exploits"], xcap
    c.num_of_device, c.min_network_size = cfg["numOfDevice"], cfg["Min_network_size"]
    c.evolve_period = cfg["evolve_period"]
    c.wl_period_base, c.wl_period_max, c.wl_cap = cfg["workload_period_base"], cfg["workload_period_max"], cfg["workload_cap"]
    c.scaling_vulnerability, c.turbo, c.zero_day = cfg["scaling_vulnerability"], cfg["turbo"], cfg["zero_day"]
    c.zero_day_mask = cfg["zero_day_mask"]
    c.att_space_n, c.def_space_n, c.default_high = cfg["att_space_n"], cfg["def_space_n"], cfg["default_high"]
    c.n_app_ids = cfg.get("n_app_ids", n_app_ids)
    c.base_line = BL.get(base_line, 4)
    c.tri_high = tri_high
    c.log_cap = int(log_cap)
    c.work_scale, c.comp_scale, c.def_scale, c.gamma = cfg["work_scale"], cfg["comp_scale"], cfg["def_scale"], cfg["gamma"]
    c.thr_p_add = D.bernoulli_threshold(cfg["p_add"])
    c.thr_p_attacker = D.bernoulli_threshold(cfg["p_attacker"])
    for i, t in enumerate(D.poisson_table(cfg["lambda_events"])):
        c.poisson_tab[i] = t
    tt = D.triangular_ceil_table(tri_mode, tri_high) + [D.M32]
    for i in range(8):
        c.tri_tab[i] = tt[i]
    c.seed = seed
    c.turbo_frac_clients = cfg.get("turbo_fraction_clients", 0.05)
    c.turbo_frac_servers = cfg.get("turbo_fraction_servers", 0.02)
    c.turbo_max_clients = cfg.get("turbo_max_clients", 200)
    c.turbo_max_servers = cfg.get("turbo_max_servers", 40)
    c.turbo_ramp_steps = cfg.get("turbo_ramp_steps", 200)
    return c


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.cyo_create.restype = C.c_void_p
        L.cyo_create.argtypes = [C.POINTER(CygConfig)] + [C.c_void_p] * 6
        L.cyo_destroy.argtypes = [C.c_void_p]
        L.cyo_set_base_line.argtypes = [C.c_void_p, C.c_int32]
        L.cyo_set_aux.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cyo_pyset_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.cyo_step.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_uint32] + [C.c_void_p] * 5 + [C.c_int]
        L.cyo_randomize.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.cyo_sample_actions.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.cyo_observe.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.cyo_philox.argtypes = [C.c_void_p] * 3
        L.cyo_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleState:
    """Canonical per-env state as numpy arrays (include/cygym_b200.h layout)."""

    def __init__(self, B, M, E, xcap, log_cap=0):
        self.B, self.M, self.E, self.xcap, self.log_cap = B, M, E, xcap, int(log_cap)
        self.logs = np.zeros((B, max(1, self.log_cap)), np.uint32)  # hop-log ring (from | to << 16), record k at k % log_cap
        self.det_slots = None                                        # [n_slots, CYG_DET_WORDS] uploaded detectors
        self.det_of_env = np.full(B, -1, np.int32)
        self.dev = np.zeros((B, M), np.uint32)
        self.ckpt = np.zeros((B, M), np.uint32)
        self.blocked = np.zeros((B, max(1, (E + 31) // 32)), np.uint32)
        self.extra = np.zeros((B, max(1, xcap)), np.uint32)
        self.scal = np.zeros((B, NSCAL), np.uint32)
        self.scal[:, 3] = 0xFFFF

    def copy(self):
        o = OracleState.__new__(OracleState)
        o.B, o.M, o.E, o.xcap, o.log_cap = self.B, self.M, self.E, self.xcap, self.log_cap
        for k in ("dev", "ckpt", "blocked", "extra", "scal", "logs", "det_of_env"):
            setattr(o, k, getattr(self, k).copy())
        o.det_slots = None if self.det_slots is None else self.det_slots.copy()
        return o

    def log_records(self, b, last=2000):
        """The last `last` hop-log records of env b in log order, int array [n, 2] of (from, to)."""
        n = int(self.scal[b, 6])
        k = min(n, last)
        if k > self.log_cap:
            raise ValueError(f"the ring keeps {self.log_cap} records, {k} are needed")
        idx = (np.arange(n - k, n) % max(1, self.log_cap)).astype(np.int64)
        r = self.logs[b, idx]
        return np.stack([r & 0xFFFF, r >> 16], axis=1).astype(np.int64)

    def set_env(self, b, st):
        """st: dict from ref_harness.extract_state()."""
        self.dev[b] = st["dev"]
        self.ckpt[b] = st["ckpt"]
        self.blocked[b, :len(st["blocked"])] = st["blocked"]
        self.extra[b] = 0
        self.extra[b, :len(st["extra"])] = st["extra"]
        self.scal[b] = st["scal"]


class Oracle:
    def __init__(self, netw, config, env_id0=0):
        self.L = lib()
        self.cfg = config
        self.M, self.E, self.X = config.M, config.E, config.X
        self.W = (self.M + 31) // 32
        self.env_id0 = env_id0
        self._keep = [np.ascontiguousarray(netw["row_ptr"], np.int32), np.ascontiguousarray(netw["col"], np.int32),
                      np.ascontiguousarray(netw["mult"], np.uint8), np.ascontiguousarray(netw["dev_static"], np.uint32),
                      np.ascontiguousarray(netw["os_val"], np.float32), np.ascontiguousarray(netw["ver_val"], np.float32)]
        assert len(self._keep[1]) == self.E
        self.h = self.L.cyo_create(C.byref(config), *[_p(a) for a in self._keep])

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.cyo_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_base_line(self, name):
        self.L.cyo_set_base_line(self.h, BL.get(name, 4))

    def new_state(self, B):
        return OracleState(B, self.M, self.E, self.cfg.xcap, self.cfg.log_cap)

    def _aux(self, st):
        self._keep_aux = (st.logs, st.det_slots, st.det_of_env)
        self.L.cyo_set_aux(self.h, _p(st.logs) if st.log_cap > 0 else None, _p(st.det_slots) if st.det_slots is not None else None,
                           _p(st.det_of_env))

    def service_detectors(self, st, seed_of_env=None):
        """Fit + upload the detector of every env whose action 10 left CYG_FL_DET_PENDING (volt:945-962 ->
        CDSimulator.py:687-695), as the product's host side does (cygym_b200/detector.py is shared: the fit is sklearn's)."""
        from cygym_b200 import detector as DET
        pend = np.nonzero(st.scal[:, 2] & 8)[0]
        for b in pend:
            model = DET.fit_detector(st.log_records(int(b)), None if seed_of_env is None else seed_of_env(int(b)))
            slot = DET.pack_detector(model)
            if st.det_of_env[b] < 0:
                st.det_of_env[b] = 0 if st.det_slots is None else len(st.det_slots)
                st.det_slots = slot[None].copy() if st.det_slots is None else np.concatenate([st.det_slots, slot[None]])
            else:
                st.det_slots[st.det_of_env[b]] = slot
            st.scal[b, 2] &= ~np.uint32(8)
        return len(pend)

    def step(self, st, hdr, mask, order=None, flags=0, n_threads=1, want_pre=False):
        """hdr [G,B,4] or [B,4] uint32; mask [G,B,W] or [B,W]."""
        hdr = np.ascontiguousarray(hdr, np.uint32)
        mask = np.ascontiguousarray(mask, np.uint32)
        if hdr.ndim == 2:
            hdr, mask = hdr[None], mask[None]
            if order is not None:
                order = order[None]
        G, B = hdr.shape[0], hdr.shape[1]
        assert B == st.B and mask.shape == (G, B, self.W)
        ostride = 0
        if order is not None:
            order = np.ascontiguousarray(order, np.uint16)
            ostride = order.shape[2]
        raw = np.zeros(B, np.float64)
        shaped = np.zeros(B, np.float64)
        done = np.zeros(B, np.int32)
        ex = np.zeros(B, np.int32)
        pre = np.zeros((B, 3, self.W), np.uint32) if want_pre else None
        self._aux(st)
        self.L.cyo_step(self.h, B, self.env_id0, _p(st.dev), _p(st.ckpt), _p(st.blocked), _p(st.extra), _p(st.scal),
                        _p(hdr), _p(mask), _p(order), ostride, G, flags, _p(raw), _p(shaped), _p(done), _p(ex), _p(pre),
                        n_threads)
        out = dict(raw=raw, shaped=shaped, done=done, exec_atype=ex)
        if want_pre:
            out["pre_masks"] = pre
        return out

    def randomize(self, st, env_mask=None):
        if env_mask is not None:
            env_mask = np.ascontiguousarray(env_mask, np.uint8)
        self.L.cyo_randomize(self.h, st.B, self.env_id0, _p(st.dev), _p(st.blocked), _p(st.extra), _p(st.scal), _p(env_mask))

    def sample_actions(self, st, mode, want_order=False):
        hdr = np.zeros((st.B, 4), np.uint32)
        mask = np.zeros((st.B, self.W), np.uint32)
        order = np.zeros((st.B, self.M), np.uint16) if want_order else None
        self.L.cyo_sample_actions(self.h, st.B, self.env_id0, _p(st.scal), mode, _p(hdr), _p(mask), _p(order), self.M)
        return (hdr, mask, order) if want_order else (hdr, mask)

    def observe(self, st, obs_mode):
        dim = 4 * self.M + self.X if obs_mode == 2 else 6 * self.M
        obs = np.zeros((st.B, dim), np.float32)
        self.L.cyo_observe(self.h, st.B, _p(st.dev), obs_mode, _p(obs))
        return obs


def pack_action(action, mode, M, order_form=False):
    """(atype, exploit_indices, device_indices, app_index) | None -> (hdr[4], mask[W], order[M] or None)."""
    W = (M + 31) // 32
    hdr = np.zeros(4, np.uint32)
    mask = np.zeros(W, np.uint32)
    order = np.zeros(M, np.uint16) if order_form else None
    m = 1 if mode in (1, "attacker") else 0
    if action is None:
        hdr[0] = ATYPE_NONE | (m << 8)
        return hdr, mask, order
    atype, ex, devs, app = action
    ex = [int(x) for x in ex][:4]
    devs = [int(d) for d in devs]
    at = max(-127, min(127, int(atype)))
    hdr[0] = (at & 0xFF) | (m << 8) | (len(ex) << 16)
    w1 = 0
    for i, x in enumerate(ex):
        w1 |= (max(-128, min(127, x)) & 0xFF) << (8 * i)
    hdr[1] = w1
    hdr[2] = len(devs)
    hdr[3] = np.uint32(int(app) & 0xFFFFFFFF)
    for d in devs:
        mask[d >> 5] |= np.uint32(1 << (d & 31))
    if order_form:
        assert len(devs) <= M
        order[:len(devs)] = devs
    else:
        assert devs == sorted(set(devs)), "mask form needs an ascending duplicate-free device list"
    return hdr, mask, order
