"""The C oracle against the UNMODIFIED reference, live (build container only: needs /root/reference).
On machines without the reference checkout the pin is the committed golden trajectories
(tests/test_oracle_golden.py), which were recorded by exactly this procedure (oracle/gen_golden.py)."""
import pytest

from oracle import ref_harness as H

pytestmark = pytest.mark.skipif(not H.available(), reason="reference checkout not present")


@pytest.mark.parametrize("kw", [
    dict(numOfDevice=10, M=20, seed=11, T=120),
    dict(numOfDevice=40, M=50, seed=12, T=80, order_form=True),
    dict(numOfDevice=25, M=35, seed=13, T=80, xcap=96, env_attrs=dict(p_add=0.5, p_attacker=0.3)),
])
def test_oracle_matches_live_reference(kw):
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import trajectory as TR
    g = TR.record(**kw)
    assert TR.replay(g, TR.OracleImpl(g), label="oracle-vs-live-reference") == len(g["kind"])


def test_product_flattening_matches_the_harness(tmp_path):
    """cygym_b200.snapshot.from_reference_env (product, duck-typed) == oracle.ref_harness.extract_* (checker),
    and the npz snapshot round-trips."""
    import numpy as np
    from cygym_b200 import snapshot
    env = H.build_env(numOfDevice=20, Max_network_size=30, seed=5)
    for t in range(30):  # move the env off its initial state
        mode = "defender" if t % 2 == 0 else "attacker"
        a = H.ref_sample_action(env, mode)
        if mode == "defender" and a[0] == 10:
            a = (8, a[1], a[2], a[3])
        H.ref_step(env, mode, a)
    env._rebuild_graph_cache()
    netw = H.extract_network(env)
    st = H.extract_state(env, netw)
    net = snapshot.from_reference_env(env)
    for k in ("row_ptr", "col", "mult", "dev_static", "os_val", "ver_val"):
        assert np.array_equal(getattr(net, k), netw[k]), k
    for k in ("dev", "ckpt", "blocked"):
        assert np.array_equal(net.template[k], st[k]), k
    sc = st["scal"].copy(); sc[1] = 0
    assert np.array_equal(net.template["scal"], sc)
    assert {k: net.cfg[k] for k in netw["cfg"] if k in net.cfg} == {k: v for k, v in netw["cfg"].items() if k in net.cfg}
    p = snapshot.save_npz(str(tmp_path / "snap.npz"), net)
    back = snapshot.load_npz(p)
    assert np.array_equal(back.col, net.col) and back.cfg == net.cfg
    assert all(np.array_equal(back.template[k], np.asarray(net.template[k], np.uint32)) for k in net.template)
