"""The DEVICE source (cygym_b200/csrc/cyg_core.cuh) compiled for the host and replayed against the
golden trajectories and against the C oracle on batched random rollouts.  This checks the logic
the CUDA kernels run on the CPU-only build container; the GPU run is tests/test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import trajectory as TR
from tests.common import (GOLDEN, GOLDEN_IDS, compare_rewards, compare_states, load_golden, oracle_for,
                          oracle_state_from_template, sanitize_actions)
from tests.emu import emu


@pytest.mark.parametrize("path", GOLDEN, ids=GOLDEN_IDS)
def test_device_source_replays_golden(path):
    g = load_golden(path)
    n = TR.replay(g, emu.EmuImpl(g), label="emu")
    assert n == len(g["kind"])


@pytest.mark.parametrize("M,subnets,B,T,kw", [
    (20, 1, 48, 120, {}),
    (50, 3, 64, 100, {}),
    (100, 8, 48, 120, {}),
    (100, 8, 32, 80, dict(p_add=0.5, p_attacker=0.4, lambda_events=1.5)),
    (33, 2, 32, 80, dict(zero_day=1, zero_day_mask=0b10)),
    (128, 4, 16, 60, {}),
    (300, 8, 160, 12, {}),                    # > 128 device slots: the 64-word generic layout
    (2000, 64, 3, 16, {}),                    # BASELINE.json config C4 shape
])
def test_device_source_matches_oracle_on_random_rollouts(M, subnets, B, T, kw):
    from cygym_b200 import synthetic_network
    net = synthetic_network(M, n_subnets=subnets, seed=M + B, **kw)
    xcap = 64
    orc, cfg = oracle_for(net, seed=99, xcap=xcap, env_id0=7)
    em = emu.Emu(dict(row_ptr=net.row_ptr, col=net.col, mult=net.mult, dev_static=net.dev_static, os_val=net.os_val,
                      ver_val=net.ver_val), cfg, env_id0=7)
    so, se = oracle_state_from_template(orc, net, B), oracle_state_from_template(orc, net, B)
    for t in range(T):
        mode = t & 1
        if t % 23 == 22:
            orc.randomize(so)
            em.randomize(se)
        ho, mo = orc.sample_actions(so, mode)
        he, me = em.sample_actions(se, mode)
        assert np.array_equal(ho, he) and np.array_equal(mo, me), f"sample_action differs at t={t}"
        ho = sanitize_actions(ho, so.scal[:, 6], mode)
        if t % 11 == 10:
            ho[::3, 0] = 0x80 | (mode << 8)  # None actions
        oo = orc.step(so, ho, mo)
        oe = em.step(se, ho, mo)
        compare_rewards(oe, oo, f"t={t}")
        compare_states(dict(dev=se.dev, ckpt=se.ckpt, blocked=se.blocked, extra=se.extra, scal=se.scal),
                       dict(dev=so.dev, ckpt=so.ckpt, blocked=so.blocked, extra=so.extra, scal=so.scal), f"t={t}")
        for om in (1, 2, 3):
            assert np.array_equal(orc.observe(so, om), em.observe(se, om))
    assert int(so.scal[:, 0].min()) == T
