"""Drive the UNMODIFIED reference step path with replayed draws (TEST INFRASTRUCTURE ONLY).

Only usable where the reference checkout exists (this container:
/root/reference).  It is how the C restatement in cyg_oracle.c is pinned and how
tests/golden/*.npz are produced (oracle/gen_golden.py); nothing here travels
to the GPU box and nothing in the product package imports it.

What it does
  * puts the stand-in modules of oracle/refshim (gym, igraph, pymetis,
    matplotlib, imageio, nashpy) and the reference directory on sys.path, and
    runs with cwd = a scratch dir that holds a synthetic CVE.csv, because
    importing volt_typhoon_env truncates ./cyberdefense_debug.log
    (volt_typhoon_env.py:26-27) and CyberDefenseSimulator() reads ./CVE.csv
    (CDSimulator.py:36);
  * replaces the names `random` and `np` INSIDE the four reference modules by
    proxies.  Outside a replay window they forward to the real generators (so
    initialize_environment() is an ordinary seeded run); inside one every RNG
    call is answered from the counter-based contract of oracle/draws.py, keyed
    by the reference file:line that made the call;
  * flattens a live reference env into the canonical struct-of-arrays layout
    (network tables + per-device words + per-env scalars) and back-compares.
"""
import contextlib
import importlib
import io
import math
import os
import random as _real_random
import sys
import tempfile

import numpy as _real_np

from . import draws as D
from . import make_cve

REF_DIR = os.environ.get("CYGYM_REFERENCE_DIR", "/root/reference")
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
_REF_MODULES = ("CDSimulatorComponents", "CDSimulator", "CyberDefenseEnv", "volt_typhoon_env")

# canonical device-word bits (mirrors include/cygym_b200.h)
DEV_COMP, DEV_KNOWN, DEV_NYA, DEV_OWNED = 1 << 0, 1 << 1, 1 << 2, 1 << 3
DEV_REMOVED, DEV_HASWL, DEV_BUSYSET, DEV_ACTSET = 1 << 4, 1 << 5, 1 << 6, 1 << 7
DEV_PT_SHIFT, DEV_BUSY_SHIFT, DEV_CBY_SHIFT = 8, 12, 20
CK_COMP, CK_KNOWN, CK_NYA, CK_REACH, CK_HASWL, CK_VALID = 1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 5, 1 << 31
ST_DC, ST_SERVER, ST_REACH = 1 << 0, 1 << 1, 1 << 2
ST_NAPPS_SHIFT, ST_VULN_SHIFT = 8, 16
FL_HAS_CKPT, FL_SETS_INIT = 1 << 0, 1 << 1
FL_DISC_SHIFT = 8
(S_STEP, S_EPOCH, S_FLAGS, S_PREV_X, S_DEF_STEP, S_ATT_STEP, S_LOGS, S_COMPCNT, S_WORK, S_DEFCOST,
 S_CLEANCOST, S_SCAN, S_REVERT, S_CKPT, S_EBLK, S_EADD) = range(16)

_SITES = {
    ("volt_typhoon_env.py", 138): D.SITE_STALL,
    ("volt_typhoon_env.py", 505): D.SITE_BLOCK,
    ("volt_typhoon_env.py", 511): D.SITE_UNBLOCK,
    ("volt_typhoon_env.py", 1136): D.SITE_ZDAY,
    ("volt_typhoon_env.py", 1189): D.SITE_PROBE,
    ("volt_typhoon_env.py", 359): D.SITE_SHUFFLE,
    ("CDSimulator.py", 298): D.SITE_WL_SAMPLE,
    ("CDSimulator.py", 308): D.SITE_WL_TRI,
    ("CDSimulator.py", 328): D.SITE_WL_LAZY,
    ("CDSimulator.py", 207): D.SITE_WL_ASSIGN,
    ("CDSimulator.py", 699): D.SITE_DETECT,
    ("CDSimulator.py", 716): D.SITE_DETECT,
    ("CyberDefenseEnv.py", 668): D.SITE_EV_POISSON,
    ("CyberDefenseEnv.py", 679): D.SITE_EV_ADD,
    ("CyberDefenseEnv.py", 675): D.SITE_EV_PICK,
    ("CyberDefenseEnv.py", 690): D.SITE_EV_ATT,
    ("CyberDefenseEnv.py", 565): D.SITE_SA_DEVS,
    ("CyberDefenseEnv.py", 566): D.SITE_SA_DEVS,
    ("CyberDefenseEnv.py", 568): D.SITE_SA_DEVS,
    ("CyberDefenseEnv.py", 567): D.SITE_SA_NDEV,
    ("CyberDefenseEnv.py", 571): D.SITE_SA_EXP,
    ("CyberDefenseEnv.py", 576): D.SITE_SA_APP,
}


def available():
    return os.path.isfile(os.path.join(REF_DIR, "volt_typhoon_env.py"))


class DrawContext:
    """The (seed, env, epoch, per-site counter) cursor shared by the proxies."""

    def __init__(self, seed=0, env_id=0, epoch=0):
        self.seed = int(seed)
        self.env_id = int(env_id)
        self.epoch = int(epoch)
        self.replay = False
        self.counts = {}
        self.trace = None  # optional list of (epoch, site, k, x)

    @contextlib.contextmanager
    def window(self):
        """One replay window == one epoch (one step / randomize / sample_action call)."""
        self.counts = {}
        self.replay = True
        try:
            yield self
        finally:
            self.replay = False
            self.epoch += 1

    def draw(self, site):
        k = self.counts.get(site, 0)
        self.counts[site] = k + 1
        x = D.draw_u32(self.seed, self.env_id, self.epoch, site, k)
        if self.trace is not None:
            self.trace.append((self.epoch, site, k, x))
        return x


_CTX = DrawContext()


def _site():
    f = sys._getframe(2)
    key = (os.path.basename(f.f_code.co_filename), f.f_lineno)
    s = _SITES.get(key)
    if s is None:
        raise RuntimeError(f"replayed RNG call from an unmapped reference site {key}")
    return s


class _RandomProxy:
    """Stands in for the `random` module inside the reference modules."""

    def __getattr__(self, name):
        return getattr(_real_random, name)

    def random(self):
        if not _CTX.replay:
            return _real_random.random()
        return _CTX.draw(_site()) / 4294967296.0

    def randint(self, a, b):
        if not _CTX.replay:
            return _real_random.randint(a, b)
        return int(a) + D.below(_CTX.draw(_site()), int(b) - int(a) + 1)

    def randrange(self, start, stop=None, step=1):
        if not _CTX.replay:
            return _real_random.randrange(start, stop, step) if stop is not None else _real_random.randrange(start)
        if stop is None:
            start, stop = 0, start
        assert step == 1
        return int(start) + D.below(_CTX.draw(_site()), int(stop) - int(start))

    def choice(self, seq):
        if not _CTX.replay:
            return _real_random.choice(seq)
        site = _site()
        if site == D.SITE_EV_PICK:
            seq = sorted(seq)  # CPython set order is not part of the contract
        if len(seq) == 0:
            raise IndexError("Cannot choose from an empty sequence")
        return seq[D.below(_CTX.draw(site), len(seq))]

    def sample(self, population, k):
        if not _CTX.replay:
            return _real_random.sample(population, k)
        site = _site()
        remaining = list(population)
        if k > len(remaining) or k < 0:
            raise ValueError("Sample larger than population or is negative")
        out = []
        for _ in range(k):
            out.append(remaining.pop(D.below(_CTX.draw(site), len(remaining))))
        return out

    def shuffle(self, lst):
        if not _CTX.replay:
            return _real_random.shuffle(lst)
        site = _site()
        remaining = list(lst)
        out = []
        while len(remaining) > 1:
            out.append(remaining.pop(D.below(_CTX.draw(site), len(remaining))))
        out.extend(remaining)
        lst[:] = out

    def uniform(self, a, b):
        if not _CTX.replay:
            return _real_random.uniform(a, b)
        raise RuntimeError("random.uniform is only reached on the dead PA path (CyberDefenseEnv.py:817)")


class _NpRandomProxy:
    def __getattr__(self, name):
        return getattr(_real_np.random, name)

    def triangular(self, left, mode, right, size=None):
        if not _CTX.replay:
            return _real_np.random.triangular(left, mode, right, size)
        assert left == 0
        site = _site()
        tab = D.triangular_ceil_table(mode, right)
        n = 1 if size is None else int(size)
        vals = [float(D.triangular_ceil_from(_CTX.draw(site), tab, right)) for _ in range(n)]
        return vals[0] if size is None else _real_np.array(vals)

    def poisson(self, lam=1.0, size=None):
        if not _CTX.replay:
            return _real_np.random.poisson(lam, size)
        assert size is None
        return D.poisson_from(_CTX.draw(_site()), D.poisson_table(lam))


class _NpProxy:
    """Stands in for `np` inside the reference modules: numpy with a proxied .random."""

    random = _NpRandomProxy()

    def __getattr__(self, name):
        return getattr(_real_np, name)


_loaded = {}


def load_reference(workdir=None):
    """Import the four reference modules behind the shims; returns the module dict."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_DIR}")
    workdir = workdir or tempfile.mkdtemp(prefix="cygym_ref_")
    make_cve.write_cve_csv(os.path.join(workdir, "CVE.csv"))
    os.chdir(workdir)
    for p in (REF_DIR, _SHIM_DIR):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_DIR)
    sys.path.insert(0, _SHIM_DIR)
    import logging
    root = logging.getLogger()
    before = list(root.handlers)
    for name in _REF_MODULES:
        _loaded[name] = importlib.import_module(name)
    # volt_typhoon_env.py:26-27 hijacks root logging at DEBUG into a file; undo.
    for h in list(root.handlers):
        if h not in before:
            root.removeHandler(h)
            try:
                h.close()
            except Exception:
                pass
    root.setLevel(logging.WARNING)
    rp, npp = _RandomProxy(), _NpProxy()
    for name in _REF_MODULES:
        mod = _loaded[name]
        if hasattr(mod, "random"):
            mod.random = rp
        if hasattr(mod, "np"):
            mod.np = npp
    import gym.spaces as _sp

    def _hook(n):
        if not _CTX.replay:
            return _real_random.randrange(n)
        return D.below(_CTX.draw(D.SITE_SA_TYPE), n)

    _sp._sample_hook = _hook
    _loaded["workdir"] = workdir
    return _loaded


def context():
    return _CTX


def seed_numpy_global(seed):
    """numpy's global stream (the REAL module, not the proxy the reference modules see): scikit-learn estimators with
    random_state=None draw from it (Detector, CDSimulator.py:683)."""
    _real_np.random.seed(int(seed))


def build_env(numOfDevice=10, Max_network_size=20, seed=1, quiet=True, **attrs):
    """initialize_environment() exactly as init_experiments.py:36-51 configures it, then
    rebuild the graph cache the way DoubleOracle.restore / reset(from_init=True) do
    (do_agent.py:891-895, volt_typhoon_env.py:1933-1936) so the attacker star is visible."""
    mods = load_reference()
    _real_random.seed(seed)
    _real_np.random.seed(seed)
    V = mods["volt_typhoon_env"].Volt_Typhoon_CyberDefenseEnv
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink if quiet else sys.stdout):
        env = V()
        env.its = 1
        env.numOfDevice = numOfDevice
        env.Max_network_size = Max_network_size
        env.base_line = "Nash"
        env.tech = "DO"
        env.mode = "defender"
        env.j_private = 1
        env.k_known = 1
        env.zero_day = False
        for k, v in attrs.items():
            setattr(env, k, v)
        env.initialize_environment()
        env._rebuild_graph_cache()
    return env


# ---------------------------------------------------------------------------
# flattening a live reference env into the canonical layout
# ---------------------------------------------------------------------------
def extract_network(env):
    """Static tables of the env's CURRENT graph cache + device statics + config."""
    net = env.simulator.subnet.net
    M = len(net)
    assert sorted(net.keys()) == list(range(M)), "device ids must be 0..M-1"
    assert M == env.Max_network_size
    row_ptr = [0]
    col, mult = [], []
    for u in range(M):
        prev = None
        for v in env._outnbrs.get(u, []):
            v = int(v)
            if prev is not None and v < prev:
                raise AssertionError("neighbour list not ascending")
            if v == prev:
                mult[-1] += 1
            else:
                col.append(v)
                mult.append(1)
            prev = v
        row_ptr.append(len(col))
    # in-neighbour lists must be the transpose (they are both read from the same igraph)
    indeg = {}
    for u in range(M):
        for v in env._innbrs.get(u, []):
            indeg[(int(v), u)] = indeg.get((int(v), u), 0) + 1
    for u in range(M):
        for e in range(row_ptr[u], row_ptr[u + 1]):
            assert indeg.get((u, col[e]), 0) == mult[e], "in/out caches disagree"
    exploits = env.simulator.exploits
    X = int(env.MaxExploits)
    assert len(exploits) <= X
    dev_static = _real_np.zeros(M, dtype=_real_np.uint32)
    os_val = _real_np.zeros(M, dtype=_real_np.float32)
    ver_val = _real_np.zeros(M, dtype=_real_np.float32)
    deg_tot = _real_np.zeros(M, dtype=_real_np.int64)
    for u in range(M):
        deg_tot[u] += sum(mult[row_ptr[u]:row_ptr[u + 1]])
        for e in range(row_ptr[u], row_ptr[u + 1]):
            deg_tot[col[e]] += mult[e]
    for i in range(M):
        d = net[i]
        w = 0
        if d.device_type == "DomainController":
            w |= ST_DC
        if d.wtype == "server":
            w |= ST_SERVER
        if d.reachable_by_attacker:
            w |= ST_REACH
        napps = len(d.apps)
        assert napps < 256
        w |= napps << ST_NAPPS_SHIFT
        for e, exp in enumerate(exploits):
            hit = any(vul.id in exp.target for app in d.apps.values() for vul in app.vulnerabilities.values())
            if hit:
                w |= 1 << (ST_VULN_SHIFT + e)
        dev_static[i] = w
        os_val[i] = env.os_to_float(d.OS)
        try:
            ver_val[i] = float(d.version)
        except Exception:
            ver_val[i] = -1.0
        assert d.anomaly_score == 0, "anomaly_score is 0 under fast_scan (volt_typhoon_env.py:46)"
    zd_mask = 0
    if env.zero_day:
        for i in (set(env.common_exploit_indices) | set(env.private_exploit_indices)):
            zd_mask |= 1 << int(i)
    cfg = dict(
        M=M, X=X, n_exploits=len(exploits), numOfDevice=int(env.numOfDevice),
        Min_network_size=int(env.Min_network_size),
        work_scale=float(env.work_scale), comp_scale=float(env.comp_scale), def_scale=float(env.def_scale),
        gamma=float(env.γ), default_high=int(env.default_high),
        lambda_events=float(env.lambda_events), p_add=float(env.p_add), p_attacker=float(env.p_attacker),
        evolve_period=int(env._evolve_period),
        workload_period_base=int(env.workload_period_base), workload_period_max=int(env.workload_period_max),
        workload_cap=(-1 if env.workload_cap is None else int(env.workload_cap)),
        scaling_vulnerability=int(bool(env.scaling_vulnerability)), turbo=int(bool(env.turbo)),
        zero_day=int(bool(env.zero_day)), zero_day_mask=zd_mask,
        att_space_n=int(env.attacker_action_space.n), def_space_n=int(env.defender_action_space.n),
        min_total_degree=int(deg_tot.min()) if M else 0,
        n_app_ids=int(env.get_num_app_indices()),
        turbo_fraction_clients=float(env.turbo_fraction_clients), turbo_fraction_servers=float(env.turbo_fraction_servers),
        turbo_max_clients=int(env.turbo_max_clients), turbo_max_servers=int(env.turbo_max_servers),
        turbo_ramp_steps=int(env.turbo_ramp_steps),
    )
    return dict(
        row_ptr=_real_np.asarray(row_ptr, dtype=_real_np.int32),
        col=_real_np.asarray(col, dtype=_real_np.int32),
        mult=_real_np.asarray(mult, dtype=_real_np.uint8),
        dev_static=dev_static, os_val=os_val, ver_val=ver_val, cfg=cfg,
    )


def _edge_index(netw):
    idx = {}
    rp, col = netw["row_ptr"], netw["col"]
    for u in range(len(rp) - 1):
        for e in range(int(rp[u]), int(rp[u + 1])):
            idx[(u, int(col[e]))] = e
    return idx


def extract_state(env, netw, epoch=0):
    """Dynamic state of a live reference env relative to the base network `netw`."""
    net = env.simulator.subnet.net
    M = netw["cfg"]["M"]
    exploits = env.simulator.exploits
    eid_of = {exp.id: i for i, exp in enumerate(exploits)}
    dev = _real_np.zeros(M, dtype=_real_np.uint32)
    ckpt = _real_np.zeros(M, dtype=_real_np.uint32)
    busyset = set()
    for d in (env._busy_devices or ()):
        busyset.add(d.id)
    has_sets = hasattr(env, "_active_ids") and hasattr(env, "_inactive_ids")
    for i in range(M):
        d = net[i]
        w = 0
        if d.isCompromised:
            w |= DEV_COMP
        if d.Known_to_attacker:
            w |= DEV_KNOWN
        if d.Not_yet_added:
            w |= DEV_NYA
        if d.attacker_owned:
            w |= DEV_OWNED
        if d.removed_before:
            w |= DEV_REMOVED
        if d.workload is not None:
            wl = d.workload
            assert not wl.adversarial, "adversarial workloads never arise on the step path"
            pt = int(wl.processing_time)
            assert 0 < pt < 8, pt
            w |= DEV_HASWL | (pt << DEV_PT_SHIFT)
        bt = d.busy_time
        assert float(bt) == int(bt) and 0 <= int(bt) < 256, bt
        w |= int(bt) << DEV_BUSY_SHIFT
        cby = 0
        for eid in d.compromised_by:
            cby |= 1 << eid_of[eid]
        w |= cby << DEV_CBY_SHIFT
        if i in busyset:
            w |= DEV_BUSYSET
        if has_sets:
            if i in env._active_ids:
                w |= DEV_ACTSET
            assert (i in env._active_ids) != (i in env._inactive_ids)
        dev[i] = w
        s = env._device_ckpts.get(i)
        if s is not None:
            c = CK_VALID
            if s["isCompromised"]:
                c |= CK_COMP
            if s["Known_to_attacker"]:
                c |= CK_KNOWN
            if s["Not_yet_added"]:
                c |= CK_NYA
            if s["reachable_by_attacker"]:
                c |= CK_REACH
            if s["workload"]:
                assert not s["workload"]["adversarial"]
                c |= CK_HASWL | (int(s["workload"]["processing_time"]) << DEV_PT_SHIFT)
            c |= int(s["busy_time"]) << DEV_BUSY_SHIFT
            cb = 0
            for eid in s["compromised_by"]:
                cb |= 1 << eid_of[eid]
            c |= cb << DEV_CBY_SHIFT
            ckpt[i] = c
    # topology relative to the base network
    eidx = _edge_index(netw)
    E = len(netw["col"])
    blocked = _real_np.zeros((E + 31) // 32, dtype=_real_np.uint32)
    extra = set()
    for u in range(M):
        for v in env._outnbrs.get(u, []):
            if (u, int(v)) not in eidx:
                extra.add((u, int(v)))
    extra_blocked = set()
    for (u, v) in env._blocked:
        e = eidx.get((int(u), int(v)))
        if e is None:
            assert (int(u), int(v)) in extra
            extra_blocked.add((int(u), int(v)))
        else:
            blocked[e >> 5] |= _real_np.uint32(1 << (e & 31))
    extra_list = sorted(extra)
    extra_words = _real_np.asarray(
        [u | (v << 12) | ((1 << 24) if (u, v) in extra_blocked else 0) for (u, v) in extra_list],
        dtype=_real_np.uint32)
    scal = _real_np.zeros(16, dtype=_real_np.uint32)
    scal[S_STEP] = env.step_num
    scal[S_EPOCH] = epoch
    fl = 0
    if env.checkpoint is not None:
        fl |= FL_HAS_CKPT
    if has_sets:
        fl |= FL_SETS_INIT
    if getattr(env.simulator.detector, "trained", False):
        fl |= 1 << 2  # CYG_FL_DET_TRAINED (CDSimulator.py:693-694)
    for i, exp in enumerate(exploits):
        if exp.discovered:
            fl |= 1 << (FL_DISC_SHIFT + i)
    scal[S_FLAGS] = fl
    prev = getattr(env, "_prev_att_potential", None)
    if prev is None:
        pn = 0xFFFF
    else:
        pn = int(round(prev * M / env.γ))
        assert abs(env.γ * pn / M - prev) < 1e-12 or abs(pn / M - prev) < 1e-12, prev
    scal[S_PREV_X] = pn | (len(extra_list) << 16)
    scal[S_DEF_STEP] = env.defender_step
    scal[S_ATT_STEP] = env.attacker_step
    scal[S_LOGS] = len(env.simulator.logger.logs)
    scal[S_COMPCNT] = env.compromised_devices_cnt
    scal[S_WORK] = env.work_done
    f32 = _real_np.array([env.defensive_cost, env.clearning_cost], dtype=_real_np.float32).view(_real_np.uint32)
    scal[S_DEFCOST], scal[S_CLEANCOST] = f32[0], f32[1]
    scal[S_SCAN] = env.scan_cnt
    scal[S_REVERT] = env.revert_count
    scal[S_CKPT] = env.checkpoint_count
    scal[S_EBLK] = env.edges_blocked
    scal[S_EADD] = env.edges_added
    return dict(dev=dev, ckpt=ckpt, blocked=blocked, extra=extra_words, scal=scal,
                defensive_cost=float(env.defensive_cost), clearning_cost=float(env.clearning_cost))


def ref_step(env, mode, action, agent_cnt=None):
    """One replayed step().  Returns (raw, shaped, done, info, pre_evolve_state)."""
    env.mode = mode
    with _CTX.window():
        state, raw, shaped, done, info, _logs = env.step(action) if agent_cnt is None else env.step(action, agent_cnt)
    return float(raw), float(shaped), bool(done), info, _real_np.asarray(state)


def ref_randomize(env):
    with _CTX.window():
        env.randomize_compromise_and_ownership()


def ref_sample_action(env, mode):
    env.mode = mode
    with _CTX.window():
        a = env.sample_action()
    return (int(a[0]), [int(x) for x in a[1]], [int(x) for x in a[2]], int(a[3]))
