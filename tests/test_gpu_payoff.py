"""Payoff-matrix evaluation (BASELINE.json config C5 in miniature): GPU rollouts vs the same loop on the
oracle, and rank-sharding invariance (two half-shards summed == one full evaluation)."""
import numpy as np
import pytest

from tests.common import oracle_for, oracle_state_from_template

pytestmark = pytest.mark.gpu


def _oracle_payoff(net, defs, atts, N, T, seed, xcap):
    from oracle import cyg_oracle as O
    from cygym_b200.vector_env import ActionBatch
    out = np.zeros((len(defs), len(atts), 10))
    for i, ds in enumerate(defs):
        for j, as_ in enumerate(atts):
            orc, _ = oracle_for(net, seed=seed, xcap=xcap, env_id0=(i * len(atts) + j) * N)
            st = oracle_state_from_template(orc, net, N)
            orc.randomize(st)
            for slot in (0, 4, 5, 8, 13, 9, 10, 12, 11):
                st.scal[:, slot] = 0
            bl = "Nash"
            dr, ar = np.zeros(N), np.zeros(N)
            for t in range(T):
                mode = t & 1
                action, nb = (ds if mode == 0 else as_).decide(t)
                if nb is not None and nb != bl:
                    bl = nb
                    orc.set_base_line(bl)
                h, m, o = ActionBatch.pack([action], mode, net.M)
                r = orc.step(st, np.repeat(h, N, 0), np.repeat(m, N, 0), n_threads=8)
                if mode == 0:
                    dr += r["raw"]
                else:
                    ar += r["raw"]
            sc = st.scal
            cols = [dr, ar, sc[:, 7], sc[:, 8], sc[:, 11], sc[:, 9].copy().view(np.float32), sc[:, 13], sc[:, 12], sc[:, 14], sc[:, 15]]
            out[i, j] = [np.asarray(c, np.float64).sum() / N for c in cols]
            out[i, j, 2] /= max(1.0, T)
    return out


def test_payoff_matrix_matches_oracle_and_shards():
    import torch
    from cygym_b200 import synthetic_network
    from cygym_b200.payoff import Strategy, evaluate_payoff_matrix, reduce_payoff
    net = synthetic_network(50, n_subnets=3, seed=21)
    defs = [Strategy(baseline_name="No Defense"), Strategy(actions=[(1, [0], [0, 3, 7, 20], 0), (6, [0], [1, 2, 3], 0), (7, [0], [5], 0)]),
            Strategy(baseline_name="Nash")]
    atts = [Strategy(baseline_name="No Attack"), Strategy(actions=[(1, [0], [], 0), (2, [0], [], 0), (1, [1], [], 0)])]
    N, T, seed, xcap = 96, 24, 5, 32
    got = evaluate_payoff_matrix(net, defs, atts, N, steps_per_episode=T, seed=seed, xcap=xcap).cpu().numpy()
    exp = _oracle_payoff(net, defs, atts, N, T, seed, xcap)
    assert got.shape == (3, 2, 10)
    assert np.allclose(got, exp, rtol=1e-5, atol=1e-6), np.abs(got - exp).max()
    # two "ranks" on one GPU: partial sums add up to the single-rank evaluation (no per-step collective needed)
    parts = [evaluate_payoff_matrix(net, defs, atts, N, steps_per_episode=T, seed=seed, xcap=xcap, rank=r, world=2, reduce=False)
             for r in range(2)]
    both = reduce_payoff(parts[0] + parts[1], N, T).cpu().numpy()
    assert np.allclose(both, got, rtol=1e-12, atol=1e-9)
    assert torch.cuda.is_available()
    # all pairs x rollouts in ONE batch (per-env base_line, one launch per turn): identical sums, any sharding
    from cygym_b200.payoff import evaluate_payoff_matrix_batched
    one = evaluate_payoff_matrix_batched(net, defs, atts, N, steps_per_episode=T, seed=seed, xcap=xcap).cpu().numpy()
    assert np.allclose(one, got, rtol=1e-12, atol=1e-9), np.abs(one - got).max()
    parts3 = [evaluate_payoff_matrix_batched(net, defs, atts, N, steps_per_episode=T, seed=seed, xcap=xcap, rank=r, world=3, reduce=False)
              for r in range(3)]
    three = reduce_payoff(parts3[0] + parts3[1] + parts3[2], N, T).cpu().numpy()
    assert np.allclose(three, got, rtol=1e-12, atol=1e-9)
    # the turns of a rollout fused 7 / 1 per launch (cyg_step_multi with one base_line row per turn): same sums
    for spl in (7, 1):
        alt = evaluate_payoff_matrix_batched(net, defs, atts, N, steps_per_episode=T, seed=seed, xcap=xcap, steps_per_launch=spl).cpu().numpy()
        assert np.allclose(alt, got, rtol=1e-12, atol=1e-9), (spl, np.abs(alt - got).max())
