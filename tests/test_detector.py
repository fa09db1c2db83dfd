"""Trained-detector support (defender actions 10 / 5): the two pieces that restate third-party behaviour are pinned
against the real thing -- the packed IsolationForest against scikit-learn's own predict, the CPython set iteration
order (the reference hands one _stall draw to every member of `flagged_senders`, a set; volt_typhoon_env.py:1062-1069)
against CPython itself."""
import ctypes as C
import random
import warnings

import numpy as np
import pytest


def _emu_and_oracle_pyset():
    from oracle import cyg_oracle as O
    from tests.emu import emu
    return emu.lib().emu_pyset_order, O.lib().cyo_pyset_order


def test_pyset_order_matches_cpython():
    fns = _emu_and_oracle_pyset()
    rng = random.Random(7)
    for trial in range(4000):
        n = rng.randint(0, 30)
        hi = rng.choice([8, 20, 33, 64, 100, 128])
        vals = [rng.randrange(hi) for _ in range(n)]
        expect = list({v for v in vals})  # a set comprehension, like the reference's
        a = np.asarray(vals, np.int32)
        for fn in fns:
            out = np.zeros(64, np.int32)
            c = fn(a.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p))
            assert list(out[:c]) == expect, (vals, expect, list(out[:c]))


def test_packed_detector_equals_sklearn_predict():
    pytest.importorskip("sklearn")
    warnings.filterwarnings("ignore")
    from cygym_b200 import detector as D
    rng = np.random.default_rng(0)
    for trial in range(6):
        M = int(rng.choice([20, 50, 100]))
        n = int(rng.integers(3, 2000))
        rec = rng.integers(0, M, size=(n, 2))
        if trial % 2 == 0:
            rec[:, 0] = rng.integers(0, 4, size=n)  # few senders, as attack logs look
        model = D.fit_detector(rec, seed=trial)
        slot = D.pack_detector(model)
        assert slot.shape == (D.DET_WORDS,)
        pts = np.asarray([(a, b) for a in range(M) for b in range(M)])
        expect = model.predict(pts) == -1
        got = np.array([D.predict_packed(slot, a, b) for a, b in pts])
        assert np.array_equal(expect, got), (trial, int((expect != got).sum()))
        # same seed, same data -> the same forest (what the parity runs rely on)
        assert np.array_equal(D.pack_detector(D.fit_detector(rec, seed=trial)), slot)
