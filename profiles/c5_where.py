"""Where the wall clock of the C5 payoff evaluation goes (one GPU): cProfile of evaluate_payoff_matrix_batched."""
import cProfile, pstats, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("C5_QUIET", "1")
from cygym_b200 import synthetic_network
from cygym_b200.payoff import Strategy, evaluate_payoff_matrix_batched
net = synthetic_network(100, n_subnets=8, seed=0)
rng = np.random.default_rng(0)
def seq(mode, L):
    out = []
    for _ in range(L):
        n = int(rng.integers(1, 40)); devs = sorted(int(d) for d in rng.choice(100, size=n, replace=False))
        at = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 11, 12, 13])) if mode == 0 else int(rng.integers(1, 3))
        out.append((at, [int(rng.integers(0, 2))], devs, int(rng.integers(0, 7))))
    return out
defs = [Strategy(baseline_name="No Defense"), Strategy(baseline_name="Preset"), Strategy(baseline_name="Nash")] + [Strategy(actions=seq(0, 8)) for _ in range(29)]
atts = [Strategy(baseline_name="No Attack"), Strategy()] + [Strategy(actions=seq(1, 8)) for _ in range(30)]
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
evaluate_payoff_matrix_batched(net, defs[:2], atts[:2], 8, steps_per_episode=4)  # warm the library
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
out = evaluate_payoff_matrix_batched(net, defs, atts, 1024, steps_per_episode=100)
torch.cuda.synchronize()
pr.disable()
print("seconds", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
