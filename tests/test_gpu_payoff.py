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


def test_parametric_strategies_in_the_batched_evaluator():
    """Actor MLPs (the DDPG best responses, do_agent.py:357-371) as strategies: the batched closed-loop evaluation equals
    the same rollout done env by env on the oracle with the reference's decode_action rule (do_agent.py:972-998,
    epsilon-free).  Both sides run the actors on the GPU, in float64 (the evaluator batches the rows of a strategy's envs
    differently from the pair-by-pair loop: in double a different cuBLAS reduction order cannot flip a `> 0` or an argmax):
    what is compared is the observation -> decode -> step path.  The reference's own actors are float32 (do_agent.py:357)."""
    import torch
    from cygym_b200 import synthetic_network
    from cygym_b200.payoff import Strategy, evaluate_payoff_matrix_batched
    from cygym_b200.vector_env import ActionBatch
    net = synthetic_network(50, n_subnets=3, seed=21)
    M, X, A = net.M, net.X, int(net.cfg.get("n_app_ids", 0))
    torch.manual_seed(3)

    class Actor(torch.nn.Module):  # the shape of do_agent.Actor
        def __init__(self, sd, ad):
            super().__init__()
            self.fc1, self.fc2, self.fc3 = torch.nn.Linear(sd, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, ad)

        def forward(self, x):
            return torch.tanh(self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x.double()))))))

    d_actor = Actor(6 * M, 14 + M + X + A).double().cuda().eval()
    a_actor = Actor(4 * M + X, 3 + M + X + A).double().cuda().eval()
    defs = [Strategy(actor=d_actor), Strategy(baseline_name="No Defense"), Strategy(actions=[(1, [0], [0, 3, 7, 20], 0), (6, [0], [1, 2, 3], 0)])]
    atts = [Strategy(actor=a_actor), Strategy(actions=[(1, [0], [], 0), (2, [0], [], 0)])]
    N, T, seed, xcap = 24, 16, 9, 32
    got = evaluate_payoff_matrix_batched(net, defs, atts, N, steps_per_episode=T, seed=seed, xcap=xcap).cpu().numpy()
    # the same on the oracle, pair by pair
    exp = np.zeros((len(defs), len(atts), 10))
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
                strat = ds if mode == 0 else as_
                if strat.actor is not None:
                    obs = torch.from_numpy(orc.observe(st, 1 + mode)).cuda()
                    with torch.no_grad():
                        raw_a = strat.actor(obs).cpu().numpy()
                    nt = 14 if mode == 0 else 3
                    acts = []
                    for b in range(N):  # decode_action, do_agent.py:972-998
                        v = raw_a[b]
                        acts.append((int(np.argmax(v[:nt])), [int(np.argmax(v[nt + M:nt + M + X]))], [int(d) for d in np.where(v[nt:nt + M] > 0)[0]],
                                     int(np.argmax(v[nt + M + X:nt + M + X + A])) if A > 0 else 0))
                    h, m, _ = ActionBatch.pack(acts, mode, M)
                else:
                    action, nb = strat.decide(t)
                    if nb is not None and nb != bl:
                        bl = nb
                        orc.set_base_line(bl)
                    h1, m1, _ = ActionBatch.pack([action], mode, M)
                    h, m = np.repeat(h1, N, 0), np.repeat(m1, N, 0)
                r = orc.step(st, h, m, n_threads=8)
                (dr if mode == 0 else ar).__iadd__(r["raw"])
            sc = st.scal
            cols = [dr, ar, sc[:, 7], sc[:, 8], sc[:, 11], sc[:, 9].copy().view(np.float32), sc[:, 13], sc[:, 12], sc[:, 14], sc[:, 15]]
            exp[i, j] = [np.asarray(c, np.float64).sum() / N for c in cols]
            exp[i, j, 2] /= max(1.0, T)
    assert np.allclose(got, exp, rtol=1e-5, atol=1e-6), np.abs(got - exp).max()
