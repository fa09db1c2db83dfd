"""Time the UNMODIFIED reference's Python step() on this machine's host cores (build container only: needs
/root/reference).  TEST INFRASTRUCTURE: writes profiles/reference_python_steps.json, which bench.py copies into
cpu_baseline.reference_python (the reference cannot travel to the GPU box; the host is stated in the record).

    python -m oracle.time_reference [--steps 2000]

Per size M in {20, 50, 100} (numOfDevice = M - 10, the reference convention of init_experiments.py:41-42 /
volt_typhoon_do.py:1473): one process on one core, then os.cpu_count() independent processes (the reference's own
mp.Pool rollout mode, do_agent.py:1737-1753).  Actions: env.sample_action() on alternating turns with the reference's
own Mersenne-Twister draws (no replay), defender action 10 rewritten to the no-op 8 (the like-for-like action set of
BASELINE.md section 3: the trained-IsolationForest branch is sklearn's arithmetic)."""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time
import warnings


def _run(args):
    M, steps, seed = args
    warnings.filterwarnings("ignore")
    from oracle import ref_harness as H
    env = H.build_env(numOfDevice=M - 10, Max_network_size=M, seed=seed)
    def one(t):
        env.mode = "defender" if t % 2 == 0 else "attacker"
        a = env.sample_action()
        if env.mode == "defender" and int(a[0]) == 10:
            a = (8, a[1], a[2], a[3])
        env.step(a)
    for t in range(50):
        one(t)
    t0 = time.perf_counter()
    for t in range(steps):
        one(t)
    return steps / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2000)
    a = ap.parse_args()
    cores = os.cpu_count() or 1
    out = {"host": platform.node(), "cpu": platform.processor() or platform.machine(), "cores": cores, "python": platform.python_version(),
           "what": "unmodified /root/reference volt_typhoon_env.Volt_Typhoon_CyberDefenseEnv.step() behind oracle/refshim stand-ins "
                   "(pure-Python igraph), sample_action() on alternating turns, defender 10 -> 8",
           "steps_per_process": a.steps, "sizes": {}}
    ctx = mp.get_context("spawn")
    for M in (20, 50, 100):
        with ctx.Pool(1) as pool:
            one = pool.map(_run, [(M, a.steps, 1)])[0]
        with ctx.Pool(cores) as pool:
            allc = sum(pool.map(_run, [(M, a.steps, 1 + i) for i in range(cores)]))
        out["sizes"][str(M)] = {"device_slots": M, "numOfDevice": M - 10, "steps_per_s_1_core": one, "steps_per_s_all_cores": allc}
        print(M, one, allc, flush=True)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "reference_python_steps.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path)


if __name__ == "__main__":
    sys.exit(main())
