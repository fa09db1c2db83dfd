"""The CPU checker (oracle/cyg_oracle.c) against the golden trajectories recorded from the
UNMODIFIED reference (oracle/gen_golden.py): this is what pins the oracle on every machine."""
import pytest

from oracle import trajectory as TR
from tests.common import GOLDEN, GOLDEN_IDS, load_golden


@pytest.mark.parametrize("path", GOLDEN, ids=GOLDEN_IDS)
def test_oracle_replays_golden(path):
    g = load_golden(path)
    n = TR.replay(g, TR.OracleImpl(g), label="oracle")
    assert n == len(g["kind"]) and n >= (30 if "c4_" in path else 100)  # the large-network files are short (O(M) reference steps)


def test_golden_cover_the_action_space():
    """Every defender type 0..13 and attacker type 0..4, grouped steps, None actions, randomize."""
    import numpy as np
    seen_def, seen_att, kinds = set(), set(), set()
    for path in GOLDEN:
        g = load_golden(path)
        for t in range(len(g["kind"])):
            kinds.add(int(g["kind"][t]))
            if int(g["kind"][t]) == TR.OP_STEP:
                (seen_att if int(g["mode"][t]) else seen_def).add(int(np.int8(g["hdr"][t][0][0] & 0xFF)))
    assert set(range(14)) <= seen_def, seen_def
    assert {0, 1, 2, 3, 4} <= seen_att, seen_att
    assert {TR.OP_STEP, TR.OP_GROUPED, TR.OP_RANDOMIZE, TR.OP_BASELINE} <= kinds
