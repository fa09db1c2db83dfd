"""Regenerate tests/golden/snapshots/ref_m30_step30.npz (build container only):  python -m oracle.gen_snapshot

A reference env (numOfDevice=20, Max_network_size=30, seed 5) is built by the UNMODIFIED reference behind the harness,
moved 30 sample_action steps off its initial state, flattened by the PRODUCT's duck-typed reader
(cygym_b200.snapshot.from_reference_env) and saved as our portable snapshot.  tests/test_gpu_parity.py loads it on the GPU
box -- where the reference does not exist -- and steps it against the oracle; tests/test_oracle_vs_reference.py checks the
flattening against the harness's own extraction where the reference does exist."""
import os
import sys
import warnings


def main():
    if os.environ.get("PYTHONHASHSEED") != "0":  # the reference's network build iterates string sets (see gen_golden.py)
        os.environ["PYTHONHASHSEED"] = "0"
        os.execv(sys.executable, [sys.executable, "-m", "oracle.gen_snapshot"])
    warnings.filterwarnings("ignore")
    from . import ref_harness as H
    from cygym_b200 import snapshot
    env = H.build_env(numOfDevice=20, Max_network_size=30, seed=5)
    for t in range(30):
        mode = "defender" if t % 2 == 0 else "attacker"
        a = H.ref_sample_action(env, mode)
        if mode == "defender" and a[0] == 10:
            a = (8, a[1], a[2], a[3])
        H.ref_step(env, mode, a)
    env._rebuild_graph_cache()
    net = snapshot.from_reference_env(env)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "snapshots")
    os.makedirs(out, exist_ok=True)
    p = snapshot.save_npz(os.path.join(out, "ref_m30_step30.npz"), net)
    print(p, os.path.getsize(p), "bytes; M =", net.M, "pairs =", len(net.col), "compromised =", int((net.template["dev"] & 1).sum()))


if __name__ == "__main__":
    main()
