"""BASELINE.json configs[1]: silver preset (logN=15) ct x ct mult + relinearize on one B200, BIT-EXACT against the
reference's own engine + CUDA kernels.  The reference package (Python + its CUDA extensions, built unmodified by
oracle/build_ref.py --engine in the build container) generates the keys and ciphertexts -- its RNG cannot be
seeded -- and both engines then run on those very tensors.  Input levels 0..14 are swept by
squaring.  Skipped when oracle/_ref/site is absent."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("reference package not installed under oracle/_ref/site")
    ref_fhe, cache = ref_engine.load()
    from liberate_b200 import fhe
    params = {k: v for k, v in fhe.params["silver"].items() if k != "devices"}
    ref = ref_fhe.ckks_engine(devices=[0], cache_folder=cache, **params)
    mine = fhe.ckks_engine(devices=[0], **params)
    assert ref.hash == mine.hash and list(ref.ctx.q) == list(mine.ctx.q)
    return ref, mine


def same(a, b):
    return all(bool((x == y).all()) for pa, pb in zip(a.data, b.data) for x, y in zip(pa, pb))


def test_mult_relin_and_rotate_bit_exact_at_every_level(engines):
    ref, mine = engines
    sk = ref.create_secret_key()
    pk = ref.create_public_key(sk)
    evk = ref.create_evk(sk)
    rotk = ref.create_rotation_key(sk, 1)
    conjk = ref.create_conjugation_key(sk)
    m = ref.example(-1, 1)
    ct = ref.encorypt(m, pk)
    # decrypt of a reference ciphertext with our engine
    assert np.abs(mine.decrode(ct, sk) - m).max() < 1e-8
    x = ct
    levels = 0
    while x.level < ref.num_levels - 1:
        r_trip = ref.cc_mult(x, x, evk, relin=False)
        m_trip = mine.cc_mult(x, x, evk, relin=False)
        assert same(r_trip, m_trip), f"triplet (lazy NTT-domain values) differs at input level {x.level}"
        r_out = ref.cc_mult(x, x, evk)
        m_out = mine.cc_mult(x, x, evk)
        assert same(r_out, m_out), f"mult+relin differs at input level {x.level}"
        assert same(ref.rotate_single(r_out, rotk), mine.rotate_single(m_out, rotk)), f"rotate differs at level {r_out.level}"
        x = r_out
        levels += 1
        if 3 <= x.level < ref.num_levels - 2:   # keep magnitudes sane (the reference's own scalar multiply: +1 level)
            x = ref.mult_scalar(x, 0.5)
    assert levels >= 7
    assert same(ref.conjugate(ct, conjk), mine.conjugate(ct, conjk))
    assert same(ref.level_up(ct, 5), mine.level_up(ct, 5))
    assert same(ref.rescale(ct), mine.rescale(ct))


def test_keys_generated_by_our_engine_work_in_the_reference(engines):
    ref, mine = engines
    sk = mine.create_secret_key()
    pk = mine.create_public_key(sk)
    evk = mine.create_evk(sk)
    m = mine.example(-1, 1)
    ct = mine.encorypt(m, pk)
    out = ref.cc_mult(ct, ct, evk)
    assert np.abs(ref.decrode(out, sk) - m * m).max() < 1e-6
    assert same(out, mine.cc_mult(ct, ct, evk))
