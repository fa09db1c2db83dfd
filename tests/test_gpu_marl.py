"""IPPO / MAPPO style rollout glue on the device (cygym_b200/marl.py) vs the reference's host-side construction
(IPPO.py:559-570) fed to the oracle's grouped step."""
import numpy as np
import pytest

from tests.common import compare_rewards, compare_states, oracle_for, oracle_state_from_template

pytestmark = pytest.mark.gpu


def test_grouped_rollout_glue_matches_reference_construction():
    import torch
    from cygym_b200 import synthetic_network
    from cygym_b200 import marl
    from cygym_b200.vector_env import ActionBatch, VectorCyberDefenseEnv
    from oracle.cyg_oracle import pack_action
    net = synthetic_network(100, n_subnets=8, seed=4)
    B, M, W = 400, net.M, net.W
    orc, _ = oracle_for(net, seed=3, xcap=32)
    env = VectorCyberDefenseEnv(net, B, seed=3, xcap=32)
    so = oracle_state_from_template(orc, net, B)
    rng = np.random.default_rng(1)
    n_types, noop = 14, 8
    for t in range(30):
        mode = t & 1
        if mode == 1:  # attacker turn: a plain random step keeps the state moving
            ho, mo = orc.sample_actions(so, 1)
            env.sample_actions(1)
            oo = orc.step(so, ho, mo)
            env.step(env.to_device(ho, mo))
            continue
        vis = marl.visibility_mask(env, "defender")
        # reference visibility: not Not_yet_added and attacker_owned (IPPO.py:74-96)
        exp_vis = ((so.dev & 4) == 0) & ((so.dev & 8) != 0)
        assert np.array_equal(vis.cpu().numpy() > 0.5, exp_vis), t
        types = rng.integers(0, n_types, size=(B, M))
        types[:, ::3] = 1                      # plenty of cleans
        full_vis = rng.random((B, M)) < 0.7   # the policy's mask is an input: exercise more than the few owned devices
        exp_idx = rng.integers(0, 6, size=B)
        app_idx = rng.integers(0, 9, size=B)
        groups = marl.grouped_actions_from_types(env, torch.from_numpy(types).to(env.device), torch.from_numpy(full_vis).to(env.device),
                                                 torch.from_numpy(exp_idx).to(env.device), torch.from_numpy(app_idx).to(env.device),
                                                 "defender", n_types, noop)
        # the reference's host-side construction, env by env (IPPO.py:559-570)
        G = n_types - 1
        hdr = np.zeros((G, B, 4), np.uint32)
        mask = np.zeros((G, B, W), np.uint32)
        for b in range(B):
            g = 0
            for ty in range(n_types):
                if ty == noop:
                    continue
                devs = [i for i in range(M) if full_vis[b, i] and types[b, i] == ty]
                if ty in marl.SINGLE_DEVICE_TYPES and devs:
                    devs = devs[:1]
                a = (ty, [int(exp_idx[b])], devs, int(app_idx[b])) if devs else (noop, [int(exp_idx[b])], [], int(app_idx[b]))
                h, m_, _ = pack_action(a, 0, M)
                hdr[g, b], mask[g, b] = h, m_
                g += 1
        for g in range(G):
            assert np.array_equal(groups[g].hdr.cpu().numpy().view(np.uint32), hdr[g]), (t, g)
            assert np.array_equal(groups[g].mask.cpu().numpy().view(np.uint32), mask[g]), (t, g)
        # the same groups from ONE kernel (cyg_group_actions), stacked
        gb = marl.grouped_actions(env, torch.from_numpy(types), None, torch.from_numpy(exp_idx), torch.from_numpy(app_idx), "defender",
                                  n_types, noop, visible=torch.from_numpy(full_vis))
        torch.cuda.synchronize()
        assert np.array_equal(gb.hdr.cpu().numpy().view(np.uint32), hdr) and np.array_equal(gb.mask.cpu().numpy().view(np.uint32), mask), t
        # ... and with the role's own visibility mask read from the bit-planes (IPPO.py:74-96)
        gr = marl.grouped_actions(env, torch.from_numpy(types), "defender", torch.from_numpy(exp_idx), torch.from_numpy(app_idx), "defender", n_types, noop)
        gt = marl.grouped_actions_from_types(env, torch.from_numpy(types).to(env.device), vis, torch.from_numpy(exp_idx).to(env.device),
                                             torch.from_numpy(app_idx).to(env.device), "defender", n_types, noop)
        torch.cuda.synchronize()
        for g in range(G):
            assert torch.equal(gr[g].hdr, gt[g].hdr) and torch.equal(gr[g].mask, gt[g].mask), (t, g)
        # type 10 with a non-empty log is the sklearn branch: the glue's caller filters it; do the same on both sides
        for g in range(G):
            bad = ((hdr[g, :, 0] & 0xFF) == 10) & (so.scal[:, 6] > 0)
            hdr[g, bad, 0] = (hdr[g, bad, 0] & ~np.uint32(0xFF)) | 8
            groups[g].hdr[:, 0] = torch.from_numpy(hdr[g, :, 0].view(np.int32)).to(env.device)
        oo = orc.step(so, hdr, mask, flags=1)
        raw, shaped, done = env.step_grouped(groups)
        torch.cuda.synchronize()
        compare_rewards(dict(raw=raw.cpu().numpy(), shaped=shaped.cpu().numpy(), done=done.cpu().numpy()), oo, f"t={t}")
    c = {k: v.cpu().numpy().view(np.uint32) for k, v in env.export_state().items()}
    compare_states(c, dict(dev=so.dev, ckpt=so.ckpt, blocked=so.blocked, extra=so.extra, scal=so.scal), "final")
