"""Host-side logic of liberate_b200 (partitioning, context scalars, engine constants) against the tables the
reference's own code produced (tests/golden/*.json from tests/golden/make_golden.py).  CPU only."""
import json

import numpy as np
import pytest

from conftest import GOLDEN


def norm(x):
    if isinstance(x, np.ndarray):
        return norm(x.tolist())
    if isinstance(x, (list, tuple)):
        return [norm(v) for v in x]
    if isinstance(x, (np.integer,)):
        return int(x)
    return x


def test_rns_partition_matches_reference():
    from liberate_b200.ntt.rns_partition import rns_partition
    cases = json.loads((GOLDEN / "partition.json").read_text())
    assert len(cases) >= 10
    for case in cases:
        L, K, D = case["args"]
        if "error" in case:
            with pytest.raises(IndexError):
                rns_partition(L, K, D)
            continue
        p = rns_partition(L, K, D)
        for name, want in case.items():
            if name == "args":
                continue
            assert norm(getattr(p, name)) == want, (case["args"], name)


def test_ckks_context_matches_reference():
    from liberate_b200.fhe.context import ckks_context
    g = json.loads((GOLDEN / "context.json").read_text())
    for case in g["contexts"]:
        c = ckks_context(**case["args"])
        assert c.q == case["q"], case["args"]
        assert c.num_scales == case["num_scales"]
        assert c.max_qbits == case["max_qbits"] and c.total_qbits == case["total_qbits"]
        assert c.generation_string == case["generation_string"]
        assert c.R_square == case["R_square"] and c.k == case["k"]
    # preset shapes stated in SURVEY.md section 8
    shapes = {"bronze": (8, 1), "silver": (17, 2), "gold": (35, 4), "platinum": (73, 6)}
    from liberate_b200.fhe.presets import params
    for name, (L, K) in shapes.items():
        kw = {k: v for k, v in params[name].items() if k != "devices"}
        c = ckks_context(**kw)
        assert (c.num_scales + 1, c.num_special_primes) == (L, K), name
