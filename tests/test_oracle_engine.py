"""The oracle's own single-device restatement of the hot path (oracle/engine_oracle.OracleEngine: rescale,
cc_mult, relinearize, key switch, rotate) reproduces the tensors of the UNMODIFIED reference engine
(tests/golden/engine_D1.json digests; inputs from engine_D1_inputs.npz).  CPU only."""
import json

import numpy as np
import pytest

from conftest import GOLDEN
from golden_utils import sha
from oracle.engine_oracle import OracleEngine


@pytest.fixture(scope="module")
def setup():
    g = json.loads((GOLDEN / "engine_D1.json").read_text())
    inp = np.load(GOLDEN / "engine_D1_inputs.npz")
    eng = OracleEngine(g["q"], g["params"]["logN"], g["params"]["num_special_primes"])
    key = lambda name: [(inp[f"{name}/{i}/0"], inp[f"{name}/{i}/1"]) for i in range(len(eng.partitions))]
    return g["digests"], inp, eng, key


def digest(x):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()


def want(d, name, poly):
    return d[name]["data"][poly][0]["sha"]


def test_rescale_and_triplet(setup):
    d, inp, eng, key = setup
    a = (inp["ct_a0"], inp["ct_a1"])
    b = (inp["ct_b0"], inp["ct_b1"])
    assert digest(eng.rescale(a[0], 0)) == want(d, "rescale_a", 0)
    assert digest(eng.rescale(a[1], 0)) == want(d, "rescale_a", 1)
    t, lv = eng.cc_mult(a, b, 0)
    assert lv == 1
    for i in range(3):
        assert digest(t[i]) == want(d, "ctt_ab", i), f"triplet d{i} (lazy representatives included)"


def test_relinearize_and_rotate(setup):
    d, inp, eng, key = setup
    a = (inp["ct_a0"], inp["ct_a1"])
    b = (inp["ct_b0"], inp["ct_b1"])
    out, lv = eng.mult(a, b, key("evk"), 0)
    assert digest(out[0]) == want(d, "ct_ab", 0) and digest(out[1]) == want(d, "ct_ab", 1)
    assert (out[0] == inp["ct_ab0"]).all() and (out[1] == inp["ct_ab1"]).all()
    r = eng.rotate(a, key("rotk1"), 1, 0)
    assert digest(r[0]) == want(d, "rot_a", 0) and digest(r[1]) == want(d, "rot_a", 1)
    r = eng.rotate(out, key("rotk1"), 1, 1)
    assert digest(r[0]) == want(d, "rot_ab", 0) and digest(r[1]) == want(d, "rot_ab", 1)
    # deeper level: square the level-1 product (partially dropped partitions)
    sq, lv2 = eng.mult(out, out, key("evk"), 1)
    assert lv2 == 2
    assert digest(sq[0]) == want(d, "square_l2", 0) and digest(sq[1]) == want(d, "square_l2", 1)
