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
