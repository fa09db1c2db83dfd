"""Golden trajectories: record from the live reference, replay against a checker (TEST INFRASTRUCTURE).

record()  -- needs /root/reference (build container only).  Drives the unmodified
             reference with replayed draws and stores, after every operation, the
             canonical state it is in.
replay()  -- needs nothing but numpy: feeds the recorded operations to any
             implementation exposing step/randomize/set_base_line on the canonical
             layout (the C oracle on CPU, the CUDA library on GPU) and compares.
"""
import json

import numpy as np

OP_STEP, OP_GROUPED, OP_RANDOMIZE, OP_BASELINE = 0, 1, 2, 3
BASELINES = ["Nash", "No Defense", "Preset", "No Attack"]
G_MAX = 4


def _policy_ops(rng, T, M, n_logs_fn, grouped_every=9, randomize_every=41, baseline_every=0, none_every=13,
                order_form=False):
    """Yield abstract ops; concrete actions are drawn when executed (they depend on live state)."""
    for t in range(T):
        if randomize_every and t % randomize_every == randomize_every - 1:
            yield ("randomize",)
        if baseline_every and t % baseline_every == baseline_every - 1:
            yield ("baseline", BASELINES[int(rng.integers(len(BASELINES)))])
        mode = "defender" if t % 2 == 0 else "attacker"
        if grouped_every and t % grouped_every == grouped_every - 1:
            yield ("grouped", mode)
        elif none_every and t % none_every == none_every - 1:
            yield ("none", mode)
        else:
            yield ("step", mode)


LOG_TAIL = 64


def _log_tail(env):
    """The last LOG_TAIL records of simulator.logger.logs as from_device | to_device << 16 (zero padded at the front)."""
    logs = env.simulator.logger.logs[-LOG_TAIL:]
    out = np.zeros(LOG_TAIL, np.uint32)
    for k, l in enumerate(logs):
        out[LOG_TAIL - len(logs) + k] = np.uint32(int(l["from_device"]) | (int(l["to_device"]) << 16))
    return out


def record(numOfDevice, M, seed, T, draw_seed=2024, order_form=False, xcap=32, grouped_every=9,
           randomize_every=41, baseline_every=0, none_every=13, env_attrs=None, env_id=0, keep_training=False, log_cap=0,
           def_types=None, att_types=None):
    """log_cap > 0: the hop log's content is part of the recording (its last LOG_TAIL records after every op) and, with
    keep_training, detector training is real: numpy's global stream -- what the reference's IsolationForest(random_state=
    None) draws from -- is seeded right before every step that trains, and the seed is recorded (`train_seed`).
    def_types: optional list the defender's sampled action TYPE is redrawn from (more scans / trainings per file)."""
    from . import cyg_oracle as O
    from . import ref_harness as H

    env = H.build_env(numOfDevice=numOfDevice, Max_network_size=M, seed=seed, **(env_attrs or {}))
    netw = H.extract_network(env)
    ctx = H.context()
    ctx.seed, ctx.env_id, ctx.epoch = draw_seed, env_id, 0
    W = (M + 31) // 32
    E = len(netw["col"])
    EW = max(1, (E + 31) // 32)
    rng = np.random.default_rng(seed * 7919 + 17)
    init = H.extract_state(env, netw, ctx.epoch)
    rec = dict(kind=[], mode=[], n_groups=[], hdr=[], mask=[], order=[], baseline=[],
               dev=[], ckpt=[], blocked=[], scal=[], extra=[], n_extra=[], raw=[], shaped=[], done=[], exec_atype=[],
               pre=[], obs_def=[], obs_att=[], sa_first=[], train_seed=[], logs_tail=[])

    def snap(raw=0.0, shaped=0.0, done=False, ex=0, state=None):
        st = H.extract_state(env, netw, ctx.epoch)
        rec["dev"].append(st["dev"]); rec["ckpt"].append(st["ckpt"])
        b = np.zeros(EW, np.uint32); b[:len(st["blocked"])] = st["blocked"]
        rec["blocked"].append(b); rec["scal"].append(st["scal"])
        x = np.zeros(xcap, np.uint32); x[:len(st["extra"])] = st["extra"]
        rec["extra"].append(x); rec["n_extra"].append(len(st["extra"]))
        rec["raw"].append(raw); rec["shaped"].append(shaped); rec["done"].append(int(done)); rec["exec_atype"].append(ex)
        pre = np.zeros((3, W), np.uint32)
        if state is not None:
            s6 = np.asarray(state).reshape(M, 6)
            for row, col in enumerate((2, 4, 5)):
                for i in range(M):
                    if s6[i, col] > 0.5:
                        pre[row, i >> 5] |= np.uint32(1 << (i & 31))
        rec["pre"].append(pre)
        rec["logs_tail"].append(_log_tail(env))
        rec["obs_def"].append(np.asarray(env._get_defender_state(), np.float32))
        rec["obs_att"].append(np.asarray(env._get_attacker_state(), np.float32))

    def push_op(kind, mode, groups, baseline=0, firsts=()):
        sf = np.full(G_MAX, -1, np.int16)
        sf[:len(firsts)] = firsts
        rec["sa_first"].append(sf)  # device_indices[0] of each sampled group in the reference's draw order
        rec["train_seed"].append(-1)
        hdr = np.zeros((G_MAX, 4), np.uint32); mask = np.zeros((G_MAX, W), np.uint32); order = np.zeros((G_MAX, M), np.uint16)
        for g, a in enumerate(groups):
            h, m, o = O.pack_action(a, mode, M, order_form)
            hdr[g], mask[g] = h, m
            if o is not None:
                order[g] = o
        rec["kind"].append(kind); rec["mode"].append(1 if mode == "attacker" else 0); rec["n_groups"].append(len(groups))
        rec["hdr"].append(hdr); rec["mask"].append(mask); rec["order"].append(order); rec["baseline"].append(baseline)

    def fix(a, mode):
        at, ex, devs, app = a
        if mode == "defender" and at == 10 and len(env.simulator.logger.logs) > 0 and not keep_training:
            at = 8  # trained-IsolationForest branch is out of the kernel's scope; under turbo the
            #         detector's predictions are never consulted (volt:1055), so training is harmless there
        if not order_form:
            devs = sorted(devs)
        return (at, ex, devs, app)

    for op in _policy_ops(rng, T, M, None, grouped_every, randomize_every, baseline_every, none_every, order_form):
        if op[0] == "randomize":
            H.ref_randomize(env)
            push_op(OP_RANDOMIZE, "defender", [])
            snap()
        elif op[0] == "baseline":
            env.base_line = op[1]
            push_op(OP_BASELINE, "defender", [], baseline=BASELINES.index(op[1]))
            snap()
        elif op[0] == "none":
            raw, shaped, done, info, state = H.ref_step(env, op[1], None)
            push_op(OP_STEP, op[1], [None])
            snap(raw, shaped, done, info["executed_atype"], state)
        elif op[0] == "step":
            a0 = H.ref_sample_action(env, op[1])
            a = fix(a0, op[1])
            if def_types and op[1] == "defender":  # a scripted type on the sampled device list (more scans / trainings)
                at = int(rng.choice(def_types))
                if at == 10 and len(env.simulator.logger.logs) > 0 and not keep_training:
                    at = 8
                a = (at, a[1], a[2], a[3])
            if att_types and op[1] == "attacker":
                a = (int(rng.choice(att_types)), a[1], a[2], a[3])
            push_op_args = (OP_STEP, op[1], [a], 0, [a0[2][0]])
            trains = op[1] == "defender" and a[0] == 10 and len(env.simulator.logger.logs) > 0 and keep_training and env.base_line == "Nash"
            tseed = 100000 + len(rec["kind"])
            if trains:
                H.seed_numpy_global(tseed)  # IsolationForest(random_state=None).fit draws from numpy's global stream
            # sample_action consumed an epoch on the reference side: record it as an explicit epoch bump
            raw, shaped, done, info, state = H.ref_step(env, op[1], a)
            push_op(*push_op_args)
            if trains:
                rec["train_seed"][-1] = tseed
            snap(raw, shaped, done, info["executed_atype"], state)
        else:  # grouped
            mode = op[1]
            ng = int(rng.integers(1, G_MAX + 1))
            groups, firsts = [], []
            for _ in range(ng):
                a0 = H.ref_sample_action(env, mode)
                firsts.append(a0[2][0])
                a = fix(a0, mode)
                at = a[0]
                if mode == "defender":
                    at = int(rng.choice([0, 1, 1, 2, 3, 8, 10, 11, 1, 5]))
                    if at == 10 and len(env.simulator.logger.logs) > 0 and not keep_training:
                        at = 1
                groups.append((at, a[1], a[2], a[3]))
            env.mode = mode
            with ctx.window():
                state, raw, shaped, done, info, _ = env.step(list(groups))
            push_op(OP_GROUPED, mode, groups, 0, firsts)
            snap(float(raw), float(shaped), bool(done), -1, state)
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["n_sample_epochs"] = np.asarray(0)
    for k in ("row_ptr", "col", "mult", "dev_static", "os_val", "ver_val"):
        out["net_" + k] = netw[k]
    for k in ("dev", "ckpt", "blocked", "extra", "scal"):
        out["init_" + k] = init[k]
    meta = dict(cfg=netw["cfg"], draw_seed=draw_seed, env_id=env_id, xcap=xcap, order_form=bool(order_form),
                numOfDevice=numOfDevice, M=M, seed=seed, T=T, log_cap=int(log_cap), scripted_def_types=bool(def_types or att_types))
    out["meta"] = np.asarray(json.dumps(meta))
    return out


class Mismatch(AssertionError):
    pass


def _check_sampled(g, t, gi, kind, order_form, got, label, scripted_def=False):
    """sample_action() parity: `got` = (hdr[4], mask[W], order[M]) an implementation sampled where the reference's own
    sample_action() produced the recorded group `gi` of op `t`.  The recorder rewrites two things afterwards (record():
    fix() turns defender 10 into 8 while the detector has logs; grouped defender steps get a scripted action type), the
    device list is stored sorted in the set form; `sa_first` keeps device_indices[0] of the draw order."""
    h, m, o = got
    rh = np.asarray(g["hdr"][t][gi], np.uint32)
    mode = int(g["mode"][t])
    sat, rat = int(h[0]) & 0xFF, int(rh[0]) & 0xFF
    if kind == OP_STEP and not scripted_def and not (sat == rat or (mode == 0 and sat == 10 and rat == 8)):
        raise Mismatch(f"{label}: op {t}: sample_action type {sat} != recorded {rat}")
    if (int(h[0]) >> 8) != (int(rh[0]) >> 8) or int(h[1]) != int(rh[1]) or int(h[3]) != int(rh[3]):
        raise Mismatch(f"{label}: op {t} group {gi}: sample_action header {[hex(int(x)) for x in h]} != {[hex(int(x)) for x in rh]}")
    nd = int(rh[2])
    if (int(h[2]) & 0xFFFF) != nd or not np.array_equal(np.asarray(m, np.uint32), np.asarray(g["mask"][t][gi], np.uint32)):
        raise Mismatch(f"{label}: op {t} group {gi}: sample_action device set differs")
    first = (int(h[2]) >> 16) - 1
    if order_form:
        ro = np.asarray(g["order"][t][gi][:nd])
        if o is not None and not np.array_equal(np.asarray(o[:nd]), ro):
            raise Mismatch(f"{label}: op {t} group {gi}: sample_action device order differs")
        if first != int(ro[0]):
            raise Mismatch(f"{label}: op {t} group {gi}: device_indices[0] {first} != {int(ro[0])}")
    elif "sa_first" in g and int(g["sa_first"][t][gi]) >= 0 and first != int(g["sa_first"][t][gi]):
        raise Mismatch(f"{label}: op {t} group {gi}: device_indices[0] {first} != {int(g['sa_first'][t][gi])}")


def replay(g, impl, check_obs=True, rtol=1e-5, label="impl", check_sample=True):
    """Feed golden trajectory `g` (dict / NpzFile) to `impl` and compare after every op.

    impl protocol (canonical numpy arrays in, canonical out):
      impl.load(init: dict(dev, ckpt, blocked, extra, scal))        # B = 1
      impl.step(hdr[G,1,4], mask[G,1,W], order[G,1,M] | None, flags) -> dict(raw, shaped, done[, exec_atype, pre_masks])
      impl.randomize();  impl.set_base_line(name);  impl.bump_epoch(n)
      impl.state() -> dict(dev[M], ckpt[M], blocked[EW], extra[xcap], scal[16])
      impl.observe(mode) -> float32[dim]
    """
    meta = json.loads(str(g["meta"]))
    M = meta["M"]
    order_form = meta["order_form"]
    has_logs = meta.get("log_cap", 0) > 0 and "logs_tail" in g
    impl.load({k: np.array(g["init_" + k]) for k in ("dev", "ckpt", "blocked", "extra", "scal")})
    T = len(g["kind"])
    for t in range(T):
        kind = int(g["kind"][t])
        if kind == OP_RANDOMIZE:
            impl.randomize()
        elif kind == OP_BASELINE:
            impl.set_base_line(BASELINES[int(g["baseline"][t])])
        else:
            G = int(g["n_groups"][t])
            hdr = np.array(g["hdr"][t][:G])[:, None, :]
            mask = np.array(g["mask"][t][:G])[:, None, :]
            order = np.array(g["order"][t][:G])[:, None, :] if order_form else None
            # the reference drew each (non-None) action with sample_action() (CyberDefenseEnv.py:555-578), one epoch per
            # group: the implementation samples too and must come up with the recorded action
            sampled = [gi for gi in range(G) if (int(hdr[gi, 0, 0]) & 0xFF) != 0x80]
            if check_sample and hasattr(impl, "sample_action"):
                for gi in sampled:
                    _check_sampled(g, t, gi, kind, order_form, impl.sample_action(int(g["mode"][t])), label, meta.get("scripted_def_types", False))
            else:
                impl.bump_epoch(len(sampled))
            out = impl.step(hdr, mask, order, 1 if kind == OP_GROUPED else 0)
            raw, shaped = float(g["raw"][t]), float(g["shaped"][t])
            tol = rtol * max(1.0, abs(raw))
            if abs(float(out["raw"][0]) - raw) > tol or abs(float(out["shaped"][0]) - shaped) > tol:
                raise Mismatch(f"{label}: op {t}: reward {out['raw'][0]}/{out['shaped'][0]} != {raw}/{shaped}")
            if int(out["done"][0]) != int(g["done"][t]):
                raise Mismatch(f"{label}: op {t}: done")
            if kind == OP_STEP and "exec_atype" in out and int(out["exec_atype"][0]) != int(g["exec_atype"][t]):
                raise Mismatch(f"{label}: op {t}: executed_atype {out['exec_atype'][0]} != {g['exec_atype'][t]}")
            if "pre_masks" in out and out["pre_masks"] is not None:
                if not np.array_equal(np.asarray(out["pre_masks"][0]), g["pre"][t]):
                    raise Mismatch(f"{label}: op {t}: pre-evolve state masks differ")
        if has_logs and int(g["train_seed"][t]) >= 0:
            impl.service_detector(int(g["train_seed"][t]))  # the reference trained inside that step (volt:945-962)
        st = impl.state()
        if has_logs and "logs_tail" in st:
            if not np.array_equal(np.asarray(st["logs_tail"], np.uint32), np.asarray(g["logs_tail"][t], np.uint32)):
                raise Mismatch(f"{label}: op {t}: hop-log records differ")
        for k in ("dev", "ckpt", "blocked"):
            if not np.array_equal(np.asarray(st[k]), g[k][t]):
                bad = np.nonzero(np.asarray(st[k]) != g[k][t])[0]
                raise Mismatch(f"{label}: op {t} kind {kind}: {k} differs at {bad[:8]}: "
                               f"{[hex(int(x)) for x in np.asarray(st[k])[bad[:8]]]} != {[hex(int(x)) for x in g[k][t][bad[:8]]]}")
        a, r = np.array(st["scal"], np.uint32), np.array(g["scal"][t], np.uint32)
        fa, fr = a[9:11].view(np.float32), r[9:11].view(np.float32)
        if not np.allclose(fa, fr, rtol=rtol, atol=1e-6):
            raise Mismatch(f"{label}: op {t}: cost counters {fa} != {fr}")
        a[9:11] = 0; r[9:11] = 0
        if not np.array_equal(a, r):
            raise Mismatch(f"{label}: op {t} kind {kind}: scalars {a} != {r}")
        nx = int(g["n_extra"][t])
        xa = sorted(int(x) & 0x1FFFFFF for x in np.asarray(st["extra"])[:nx])
        xr = sorted(int(x) for x in g["extra"][t][:nx])
        if xa != xr:
            raise Mismatch(f"{label}: op {t}: extra edges {xa} != {xr}")
        if check_obs:
            for mode, key in ((1, "obs_def"), (2, "obs_att")):
                o = np.asarray(impl.observe(mode))
                if not np.array_equal(o, g[key][t]):
                    raise Mismatch(f"{label}: op {t}: observation mode {mode} differs")
    return T


def log_tail_of_ring(ring, n, cap):
    """The last LOG_TAIL records of a hop-log ring (record k at k % cap), zero padded at the front."""
    k = min(n, LOG_TAIL, cap)
    out = np.zeros(LOG_TAIL, np.uint32)
    if k:
        out[LOG_TAIL - k:] = np.asarray(ring, np.uint32)[(np.arange(n - k, n) % cap).astype(np.int64)]
    return out


class OracleImpl:
    """The replay() protocol on top of the C oracle (B = 1)."""

    def __init__(self, g, base_line="Nash"):
        from . import cyg_oracle as O
        meta = json.loads(str(g["meta"]))
        self.meta = meta
        netw = {k: np.array(g["net_" + k]) for k in ("row_ptr", "col", "mult", "dev_static", "os_val", "ver_val")}
        self.cfg = O.make_config(meta["cfg"], len(netw["col"]), seed=meta["draw_seed"], xcap=meta["xcap"], base_line=base_line,
                                 log_cap=meta.get("log_cap", 0))
        self.orc = O.Oracle(netw, self.cfg, env_id0=meta["env_id"])
        self.st = self.orc.new_state(1)

    def load(self, init):
        self.st.dev[0] = init["dev"]; self.st.ckpt[0] = init["ckpt"]
        self.st.blocked[0] = 0; self.st.blocked[0, :len(init["blocked"])] = init["blocked"]
        self.st.extra[0] = 0; self.st.extra[0, :len(init["extra"])] = init["extra"]
        self.st.scal[0] = init["scal"]

    def bump_epoch(self, n):
        self.st.scal[0, 1] += np.uint32(n)

    def sample_action(self, mode):
        h, m, o = self.orc.sample_actions(self.st, mode, want_order=True)
        return h[0], m[0], o[0]

    def step(self, hdr, mask, order, flags):
        return self.orc.step(self.st, hdr, mask, order, flags=flags, want_pre=True)

    def randomize(self):
        self.orc.randomize(self.st)

    def set_base_line(self, name):
        self.orc.set_base_line(name)

    def state(self):
        d = dict(dev=self.st.dev[0], ckpt=self.st.ckpt[0], blocked=self.st.blocked[0], extra=self.st.extra[0], scal=self.st.scal[0])
        if self.st.log_cap > 0:
            d["logs_tail"] = log_tail_of_ring(self.st.logs[0], int(self.st.scal[0, 6]), self.st.log_cap)
        return d

    def service_detector(self, seed):
        self.orc.service_detectors(self.st, lambda b: seed)

    def observe(self, mode):
        return self.orc.observe(self.st, mode)[0]
