"""Batched payoff-matrix evaluation: the GPU counterpart of DoubleOracle.build_payoff_matrices /
simulate_game (do_agent.py:1666-1870, :1875-2089) for the non-neural Strategy flavours
(strategy.py:25-60): a baseline name, a fixed action sequence, or the no-op.

Every (defender strategy i, attacker strategy j) pair is rolled out `n_rollouts` times for
`steps_per_episode` alternating turns (defender on even t, do_agent.py:2053).  Rollouts are
independent, so they shard over ranks with no per-step collective: rank r runs rollouts
[r*N/world, (r+1)*N/world) of every pair and holds a partial [n_def, n_att, 10] sum; ONE
all-reduce(sum) of that tensor (NCCL over NVLink on GPUs, gloo in the CPU tests) finishes the
evaluation.  The 10 columns are simulate_game's return tuple (do_agent.py:2078-2089).
"""
import numpy as np

COLUMNS = ("defender_return", "attacker_return", "compromised_fraction", "jobs_completed", "scan_count",
           "defensive_cost", "checkpoint_count", "revert_count", "edges_blocked", "edges_added")


class Strategy:
    """The non-neural flavours of strategy.Strategy: baseline_name | actions (fixed sequence) | neither (no-op)."""

    def __init__(self, baseline_name=None, actions=None):
        self.baseline_name = baseline_name
        self.actions = list(actions) if actions is not None else None

    def decide(self, t_step):
        """(action, base_line to set or None) -- _strategy_decide_action (do_agent.py:707-764)."""
        if self.baseline_name is not None:
            return None, self.baseline_name
        if self.actions:
            return self.actions[t_step % len(self.actions)], None
        return None, "Nash"


def shard_range(total, rank, world):
    """Contiguous split of `total` items over `world` ranks (the first total % world ranks get one more)."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_payoff(partial_sums, n_rollouts, steps_per_episode, group=None):
    """All-reduce the per-rank [n_def, n_att, 10] sums and turn them into simulate_game's averages."""
    import torch
    import torch.distributed as dist
    t = partial_sums.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = t / float(n_rollouts)
    out[..., 2] = out[..., 2] / max(1.0, float(steps_per_episode))  # avg_compromised_fraction (do_agent.py:2076)
    return out


def evaluate_payoff_matrix(network, def_strategies, att_strategies, n_rollouts, steps_per_episode=100, seed=0,
                           device="cuda:0", rank=0, world=1, xcap=16, group=None, reduce=True):
    """Returns a float64 tensor [n_def, n_att, 10] (averages) on `device`."""
    import torch
    from .vector_env import ActionBatch, VectorCyberDefenseEnv
    from . import _capi as K
    lo, hi = shard_range(n_rollouts, rank, world)
    nloc = hi - lo
    nd, na = len(def_strategies), len(att_strategies)
    sums = torch.zeros(nd, na, len(COLUMNS), dtype=torch.float64, device=device)
    if nloc > 0:
        M = network.M
        for i, ds in enumerate(def_strategies):
            for j, as_ in enumerate(att_strategies):
                # env ids are unique per (pair, rollout) so that every rollout has its own draw stream
                env = VectorCyberDefenseEnv(network, nloc, device=device, seed=seed,
                                            env_id0=(i * na + j) * n_rollouts + lo, xcap=xcap)
                env.randomize_compromise_and_ownership()
                # counters the reference zeroes before a rollout (do_agent.py:2038-2045)
                s = env.scalars
                for slot in (K.S_STEP, K.S_DEF_STEP, K.S_ATT_STEP, K.S_WORK, K.S_CKPT, K.S_DEFCOST, K.S_CLEANCOST,
                             K.S_REVERT, K.S_SCAN):
                    s[:, slot] = 0
                def_r = torch.zeros(nloc, dtype=torch.float64, device=device)
                att_r = torch.zeros(nloc, dtype=torch.float64, device=device)
                for t in range(steps_per_episode):
                    mode = t & 1
                    action, bl = (ds if mode == 0 else as_).decide(t)
                    if bl is not None and bl != env.base_line:
                        env.set_base_line(bl)
                    hdr, mask, order = ActionBatch.pack([action], mode, M, order_form=False) if (
                        action is None or list(action[2]) == sorted(set(action[2]))) else ActionBatch.pack([action], mode, M, order_form=True)
                    ab = env.to_device(np.repeat(hdr, nloc, 0), np.repeat(mask, nloc, 0),
                                       None if order is None else np.repeat(order, nloc, 0))
                    raw, _, _ = env.step(ab)
                    if mode == 0:
                        def_r += raw.double()
                    else:
                        att_r += raw.double()
                info = env.info()
                cols = [def_r, att_r, info["Compromised_devices"].double(), info["work_done"].double(),
                        info["Scan_count"].double(), info["defensive_cost"].double(), info["checkpoint_count"].double(),
                        info["revert_count"].double(), info["Edges Blocked"].double(), info["Edges Added"].double()]
                sums[i, j] = torch.stack([c.sum() for c in cols])
                env.close()
    if not reduce:
        return sums
    return reduce_payoff(sums, n_rollouts, steps_per_episode, group=group)


def evaluate_payoff_matrix_batched(network, def_strategies, att_strategies, n_rollouts, steps_per_episode=100, seed=0,
                                   device="cuda:0", rank=0, world=1, xcap=16, group=None, reduce=True, steps_per_launch=10):
    """Same result as evaluate_payoff_matrix(), with ALL (pair, rollout) combinations in one batch: the flattened
    index g = pair * n_rollouts + rollout is split contiguously over the ranks (env id == g, so the draw streams
    do not depend on the number of ranks), `steps_per_launch` turns are ONE kernel launch (cyg_step_multi: the records
    stay in shared memory between the turns), baselines act through per-env base_line rows (one per turn of the
    launch, cyg_set_base_line_per_env_steps) and fixed-sequence strategies through per-pair action rows gathered to
    the envs on the device.  Strategies with unsorted / repeated device lists are not handled here (use the
    per-pair evaluator)."""
    import torch
    from .vector_env import ActionBatch, VectorCyberDefenseEnv
    from . import _capi as K
    nd, na = len(def_strategies), len(att_strategies)
    P = nd * na
    lo, hi = shard_range(P * n_rollouts, rank, world)
    nloc = hi - lo
    sums = torch.zeros(P, len(COLUMNS), dtype=torch.float64, device=device)
    if nloc > 0:
        M = network.M
        env = VectorCyberDefenseEnv(network, nloc, device=device, seed=seed, env_id0=lo, xcap=xcap)
        pair_of_env = (torch.arange(lo, hi, device=device) // n_rollouts)
        i_of_pair = torch.arange(P, device=device) // na
        j_of_pair = torch.arange(P, device=device) % na
        env.randomize_compromise_and_ownership()
        s = env.scalars
        for slot in (K.S_STEP, K.S_DEF_STEP, K.S_ATT_STEP, K.S_WORK, K.S_CKPT, K.S_DEFCOST, K.S_CLEANCOST, K.S_REVERT, K.S_SCAN):
            s[:, slot] = 0
        bl_pair = torch.full((P,), K.BASE_LINES["Nash"], dtype=torch.uint8, device=device)
        ret = torch.zeros(2, nloc, dtype=torch.float64, device=device)
        fuse = max(1, int(steps_per_launch)) if network.W <= 4 else 1
        t0 = 0
        while t0 < steps_per_episode:
            # one chunk of turns = ONE launch (cyg_step_multi): per-pair action rows and base_line codes of every turn
            # of the chunk are packed on the host ([Tc, P, ..], tiny) and gathered to the envs on the device
            Tc = min(fuse, steps_per_episode - t0)
            hdr_p, mask_p, bl_p = [], [], []
            for t in range(t0, t0 + Tc):
                mode = t & 1
                strategies = def_strategies if mode == 0 else att_strategies
                decided = [st.decide(t) for st in strategies]
                for a, _ in decided:
                    if a is not None and list(a[2]) != sorted(set(int(d) for d in a[2])):
                        raise NotImplementedError("unsorted device_indices: use evaluate_payoff_matrix()")
                hdr, mask, _ = ActionBatch.pack([a for a, _ in decided], mode, M)
                which = i_of_pair if mode == 0 else j_of_pair
                new_bl = torch.tensor([K.BASE_LINES.get(b, 4) if b is not None else 255 for _, b in decided], dtype=torch.uint8, device=device)[which]
                bl_pair = torch.where(new_bl == 255, bl_pair, new_bl)   # a strategy that sets no base_line leaves it as it was
                hdr_p.append(torch.from_numpy(hdr.view(np.int32)).to(device)[which])
                mask_p.append(torch.from_numpy(mask.view(np.int32)).to(device)[which])
                bl_p.append(bl_pair)
            if Tc == 1:
                env.set_base_line_per_env(bl_p[0][pair_of_env])
                raw, _, _ = env.step(ActionBatch(hdr_p[0][pair_of_env].contiguous(), mask_p[0][pair_of_env].contiguous()))
                ret[t0 & 1] += raw.double()
            else:
                env.set_base_line_per_env(torch.stack(bl_p)[:, pair_of_env])
                raw, _, _ = env.step_many(torch.stack(hdr_p)[:, pair_of_env].contiguous(), torch.stack(mask_p)[:, pair_of_env].contiguous())
                r64 = raw.double()
                ret[t0 & 1] += r64[0::2].sum(0)
                ret[(t0 + 1) & 1] += r64[1::2].sum(0)
            t0 += Tc
        info = env.info()
        cols = torch.stack([ret[0], ret[1], info["Compromised_devices"].double(), info["work_done"].double(),
                            info["Scan_count"].double(), info["defensive_cost"].double(), info["checkpoint_count"].double(),
                            info["revert_count"].double(), info["Edges Blocked"].double(), info["Edges Added"].double()], dim=1)
        sums.index_add_(0, pair_of_env, cols)
        env.close()
    sums = sums.view(nd, na, len(COLUMNS))
    if not reduce:
        return sums
    return reduce_payoff(sums, n_rollouts, steps_per_episode, group=group)
