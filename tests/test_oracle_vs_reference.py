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


def test_reference_pickle_round_trip_through_the_kernels_state(tmp_path):
    """The reference's on-disk format both ways (init_experiments.py:53-62, volt:1904-1925): a pickled env is loaded into
    the struct-of-arrays form, stepped THERE (device source, host build), written back into the reference's object graph
    and pickled; the reloaded reference env and our state then keep stepping in lock step under replayed draws."""
    import pickle
    import numpy as np
    from cygym_b200 import snapshot
    from oracle import cyg_oracle as O
    from oracle import trajectory as TR
    from tests.emu import emu
    env = H.build_env(numOfDevice=20, Max_network_size=30, seed=9)
    p0 = str(tmp_path / "initial_net_DO_its1.pkl")
    with open(p0, "wb") as f:
        pickle.dump(env, f)   # what init_experiments.py writes
    net, env2 = snapshot.load_reference_pickle(p0)
    ctx = H.context()
    ctx.seed, ctx.env_id, ctx.epoch = 4242, 3, 0
    cfg = O.make_config(net.cfg, net.E, seed=4242, xcap=64)
    em = emu.Emu(dict(row_ptr=net.row_ptr, col=net.col, mult=net.mult, dev_static=net.dev_static, os_val=net.os_val, ver_val=net.ver_val),
                 cfg, env_id0=3)
    st = O.OracleState(1, net.M, net.E, 64)
    st.set_env(0, dict(dev=net.template["dev"], ckpt=net.template["ckpt"], blocked=net.template["blocked"], extra=net.template["extra"],
                       scal=net.template["scal"]))

    def ours_step(mode, a):
        h, m, o = O.pack_action(a, mode, net.M, order_form=True)
        return em.step(st, h[None, None], m[None, None], o[None, None])

    def sample(e_ref, mode):
        a = H.ref_sample_action(e_ref, mode)
        st.scal[0, 1] += 1  # our side replays the sampled action: the epoch sample_action() consumed
        return (8 if mode == "defender" and a[0] == 10 else a[0], a[1], a[2], a[3])

    # phase 1: OUR side alone moves on (the reference env2 only provides the sampled actions' draw epochs)
    acts = []
    for t in range(40):
        mode = "defender" if t % 2 == 0 else "attacker"
        a = sample(env2, mode)
        acts.append((mode, a))
        ours_step(mode, a)
        ctx.epoch += 1       # the step epoch the reference did not take
    # write our state into the reference object graph, pickle, reload
    state = dict(dev=st.dev[0], ckpt=st.ckpt[0], blocked=st.blocked[0], extra=st.extra[0], scal=st.scal[0])
    p1 = snapshot.write_reference_pickle(str(tmp_path / "after_40_steps.pkl"), env2, net, state)
    with open(p1, "rb") as f:
        env3 = pickle.load(f)
    blocked_before = set(env3._blocked)
    netw = H.extract_network(env3)  # the base graph now includes the extra edges our side added
    st3 = H.extract_state(env3, netw, ctx.epoch)
    # same dynamic state, seen through the harness' own flattening (device words, checkpoints, scalars)
    assert np.array_equal(st3["dev"], st.dev[0]) and np.array_equal(st3["ckpt"], st.ckpt[0])
    a3, b3 = st3["scal"].copy(), st.scal[0].copy()
    a3[1] = b3[1] = 0
    a3[3] &= 0xFFFF; b3[3] &= 0xFFFF  # the extra-edge count: the reloaded graph holds them as base edges
    assert np.array_equal(a3, b3), (a3, b3)
    assert len(blocked_before) == int(sum(bin(int(x)).count("1") for x in st.blocked[0])) + sum(1 for x in st.extra[0][: int(st.scal[0, 3]) >> 16] if int(x) & (1 << 24))
    # phase 2: both go on, in lock step, from the written-back state
    for t in range(40, 70):
        mode = "defender" if t % 2 == 0 else "attacker"
        a = sample(env3, mode)
        raw, shaped, done, info, _ = H.ref_step(env3, mode, a)
        o = ours_step(mode, a)
        assert abs(float(o["raw"][0]) - raw) <= 1e-5 * max(1.0, abs(raw)), t
    st4 = H.extract_state(env3, netw, ctx.epoch)
    assert np.array_equal(st4["dev"], st.dev[0])
    assert int(st4["scal"][7]) == int(st.scal[0, 7]) and int(st4["scal"][6]) == int(st.scal[0, 6])
