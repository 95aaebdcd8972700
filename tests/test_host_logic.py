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


def test_base_prime_is_a_partition_of_its_own_and_scale_tables_straddle_2_42_only_at_42_bits():
    """fhe/executor.LevelPlan chooses the FP64 key-switch pipeline unless a MULTI-limb partition holds a prime >= 2^42 while
    FP64 target limbs exist.  (1) the base prime (60 bits) is always a partition of its own (part.py:29-34; golden sweep + a
    denser sweep), so with scale primes below 2^42 the only wide partition has one limb; (2) a scale_bits table lies
    entirely below 2^42 (bits <= 41) or entirely above (bits >= 43) -- only the 42-bit tables straddle the limit, which is
    the one case that takes the integer pipeline (tests/test_gpu_engine.py: _sb42)."""
    import json
    from conftest import GOLDEN
    from liberate_b200.ntt.rns_partition import rns_partition
    cases = [tuple(it["args"]) for it in json.loads((GOLDEN / "partition.json").read_text()) if "error" not in it]
    cases += [(L, K, D) for L in range(4, 40, 3) for K in (1, 2, 3, 4, 6) for D in (1, 2, 4, 8) if L > K * D]
    for L, K, D in cases:
        try:
            p = rns_partition(L, K, D)
        except IndexError:
            continue
        own = [part for part in p.partitions if p.base_prime_idx in part]
        assert own == [[p.base_prime_idx]], (L, K, D, own)
    from liberate_b200.fhe.context.ckks_context import tables
    t = tables()
    for key, primes in t["scale_primes"].items():
        if not primes:
            continue
        bits = int(key.split(",")[0])
        n_wide = sum(q >= (1 << 42) for q in primes)
        if bits <= 41:
            assert n_wide == 0, key
        elif bits >= 43:
            assert n_wide == len(primes), key
        else:
            assert 0 < n_wide < len(primes), key
    assert all(q >= (1 << 42) for per_n in t["message_special_primes"]["60"].values() for q in per_n)


def test_ntt_domain_galois_permutation_matches_the_oracle_transform():
    """engine.rotate_hoisted moves rotation keys to their NTT-domain pre-image with galois_ntt_index: NTT(pi_g(x))[i] ==
    NTT(x)[P[i]].  Checked against the oracle's forward transform (kern.cu:236-275 restated) and a schoolbook Galois map."""
    import numpy as np
    from conftest import primes_for
    from liberate_b200.fhe.ckks_engine import galois_ntt_index
    from oracle import oracle as O
    logN = 12
    N = 1 << logN
    P = O.Params(primes_for(logN, 1, 1), logN)
    q = np.array(P.q, dtype=np.int64)
    rng = np.random.default_rng(3)
    x = rng.integers(0, q[:, None], (len(q), N), dtype=np.int64)

    def ntt(a):
        a = a.copy()
        O.C.ntt(a, P.psi, P._2q, *P.mont)
        O.C.reduce_2q(a, P._2q)
        return a
    A = ntt(x)
    for g in (3, pow(3, 5, 2 * N), 2 * N - 1, pow(3, -7, 2 * N)):
        j = np.arange(N)
        t = (g * j) % (2 * N)
        moved = np.zeros_like(x)
        moved[:, t % N] = np.where(t < N, x, (-x) % q[:, None])
        idx = galois_ntt_index(g, logN).numpy()
        assert (ntt(moved) == A[:, idx]).all(), g
        inv = galois_ntt_index(pow(g, -1, 2 * N), logN).numpy()
        assert (idx[inv] == j).all()
