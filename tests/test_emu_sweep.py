"""Randomised sweep of the DEVICE source (host build) against the C oracle over network shapes and env
attributes the goldens do not cover one by one: dense multi-edges, workload caps, non-scaled arrivals, odd
evolve periods, zero-day sets, high attacker-arrival rates, all four base lines, grouped and order-form steps."""
import numpy as np
import pytest

from tests.common import compare_rewards, compare_states, oracle_for, oracle_state_from_template, sanitize_actions
from tests.emu import emu

CASES = [
    dict(M=24, sub=2, kw=dict(workload_cap=3, lambda_events=2.0, p_add=0.6)),
    dict(M=40, sub=1, kw=dict(scaling_vulnerability=0, evolve_period=3)),
    dict(M=64, sub=4, kw=dict(zero_day=1, zero_day_mask=0b11, p_attacker=0.5, p_add=0.5, lambda_events=1.2)),
    dict(M=96, sub=8, kw=dict(comp_scale=50.0, def_scale=0.3, work_scale=2.5, default_high=5)),
    dict(M=100, sub=8, kw=dict(workload_period_base=3, workload_period_max=12)),       # arrivals fire often
    dict(M=128, sub=2, kw=dict(evolve_period=1, lambda_events=3.0, p_add=0.3)),
    dict(M=70, sub=3, kw=dict(workload_cap=0)),
    dict(M=160, sub=5, kw=dict(p_attacker=0.3, p_add=0.4, lambda_events=1.0)),        # 64-word layout + extra edges
]


def _dense_multi(net, rng):
    """Turn ~10 % of the pairs into multi-edges (multiplicity 2..4)."""
    m = net.mult.copy()
    idx = rng.choice(len(m), size=max(1, len(m) // 10), replace=False)
    m[idx] = rng.integers(2, 5, size=len(idx))
    net.mult = m
    return net


@pytest.mark.parametrize("case", CASES, ids=[f"M{c['M']}" for c in CASES])
def test_sweep(case):
    from cygym_b200 import synthetic_network
    rng = np.random.default_rng(case["M"])
    net = _dense_multi(synthetic_network(case["M"], n_subnets=case["sub"], seed=case["M"] * 3 + 1, **case["kw"]), rng)
    _sweep(net, rng)


def test_snapshot_of_a_reference_env():
    """The committed snapshot of a reference env 30 steps into an episode (oracle/gen_snapshot.py, flattened by
    cygym_b200.snapshot.from_reference_env) loads without the reference and steps from that state -- device source and
    oracle agree (the GPU twin is tests/test_gpu_parity.py::test_snapshot_of_a_reference_env_steps_like_the_oracle)."""
    import os
    from cygym_b200 import snapshot
    from tests.common import ROOT
    net = snapshot.load_npz(os.path.join(ROOT, "tests", "golden", "snapshots", "ref_m30_step30.npz"))
    assert net.M == 30 and int(net.template["scal"][0]) == 30 and int((net.template["dev"] & 1).sum()) > 0
    _sweep(net, np.random.default_rng(30))


def _sweep(net, rng):
    xcap, B, T = 200, 40, 70
    W = net.W
    for base_line in ("Nash", "No Defense", "Preset", "No Attack"):
        orc, cfg = oracle_for(net, seed=31, xcap=xcap, env_id0=11, base_line=base_line)
        em = emu.Emu(dict(row_ptr=net.row_ptr, col=net.col, mult=net.mult, dev_static=net.dev_static, os_val=net.os_val,
                          ver_val=net.ver_val), cfg, env_id0=11)
        so, se = oracle_state_from_template(orc, net, B), oracle_state_from_template(orc, net, B)
        steps = T if base_line == "Nash" else 12
        for t in range(steps):
            mode = t & 1
            if t % 17 == 16:
                orc.randomize(so)
                em.randomize(se)
            ho, mo = orc.sample_actions(so, mode)
            he, me = em.sample_actions(se, mode)
            assert np.array_equal(ho, he) and np.array_equal(mo, me)
            ho = sanitize_actions(ho, so.scal[:, 6], mode)
            if t % 9 == 8:
                ho[::4, 0] = 0x80 | (mode << 8)
            order, flags = None, 0
            if t % 7 == 6:  # explicit (shuffled, possibly repeated) device_indices order
                order = np.zeros((B, net.M), np.uint16)
                for b in range(B):
                    devs = [d for d in range(net.M) if (mo[b, d >> 5] >> (d & 31)) & 1]
                    rng.shuffle(devs)
                    n = int(ho[b, 2]) & 0xFFFF  # the high half carries device_indices[0] of the draw
                    devs = (devs + devs)[:n] if n > len(devs) else devs[:n]
                    order[b, :len(devs)] = devs
            if t % 13 == 12:
                flags = 2  # agent_cnt != len(net): work / arrivals / counters skipped (volt:1207)
            if t % 10 == 9 and order is None:  # grouped step: three groups
                h3 = np.stack([ho, np.roll(ho, 1, 0), np.roll(ho, 2, 0)])
                m3 = np.stack([mo, np.roll(mo, 1, 0), np.roll(mo, 2, 0)])
                h3[:, :, 0] = (h3[:, :, 0] & ~np.uint32(0x100)) | np.uint32(mode << 8)
                bad = ((h3[:, :, 0] & 0xFF) == 10) | ((h3[:, :, 0] & 0xFF) == 11) & (h3[:, :, 2] == 0)
                h3[:, :, 0] = np.where(bad, (h3[:, :, 0] & ~np.uint32(0xFF)) | 8, h3[:, :, 0])
                oo = orc.step(so, h3, m3, flags=1)
                oe = em.step(se, h3, m3, flags=1)
            else:
                oo = orc.step(so, ho, mo, order, flags=flags)
                oe = em.step(se, ho, mo, order, flags=flags)
            compare_rewards(oe, oo, f"{base_line} t={t}")
            compare_states(dict(dev=se.dev, ckpt=se.ckpt, blocked=se.blocked, extra=se.extra, scal=se.scal),
                           dict(dev=so.dev, ckpt=so.ckpt, blocked=so.blocked, extra=so.extra, scal=so.scal), f"{base_line} t={t}")
            if t % 5 == 0:
                for om in (1, 2, 3):
                    assert np.array_equal(orc.observe(so, om), em.observe(se, om))
