"""The sampler oracle (oracle/csprng_oracle.py) is pinned by the RFC 8439 block-function vector and by outputs of the
reference's own Python code (tests/golden/csprng.npz, made by tests/golden/make_golden_csprng.py)."""
import numpy as np

from conftest import GOLDEN
from oracle import csprng_oracle as R


def test_chacha20_block_rfc8439_vector():
    # RFC 8439 section 2.3.2: key 00..1f, block counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00.
    # In the 64-bit-counter layout used here words 12..15 are counter_lo, counter_hi, nonce0, nonce1.
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    words = R.chacha20_block(key + [0x4A000000, 0x00000000], [(0x09000000 << 32) | 1])[0]
    want = [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
            0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    assert [int(w) for w in words] == want


def test_chacha20_matches_reference_python():
    g = np.load(GOLDEN / "csprng.npz")
    got = R.chacha20_states(g["states"].astype(np.uint64))
    assert (got.astype(np.int64) == g["chacha20"]).all()


def test_cdt_tree_matches_reference_builder():
    g = np.load(GOLDEN / "csprng.npz")
    tree, size, depth = R.build_cdt_tree(128, 3.2)
    assert size == int(g["tree_size"]) and depth == int(g["tree_depth"])
    assert (tree == g["tree"]).all()


def test_samplers_are_sane():
    rng = np.random.default_rng(3)
    key = [int(x) for x in rng.integers(0, 1 << 32, 10)]
    words = R.chacha20_block(key, range(1000, 1000 + 4096))
    q = 1152921504606830593
    u = R.randint(words, q)
    assert u.min() >= 0 and u.max() < q and abs(u.mean() / q - 0.5) < 0.02
    t = R.randint(words, 3, shift=-1)
    assert set(np.unique(t)) == {-1, 0, 1}
    tree, size, depth = R.build_cdt_tree()
    gsn = R.discrete_gaussian(words, tree, size, depth)
    assert abs(gsn.std() - 3.2) < 0.1 and abs(gsn.mean()) < 0.1 and np.abs(gsn).max() < 32
    x = rng.uniform(-5, 5, 1000)
    r = R.randround(x, R.chacha20_block(key, range(63)))
    assert np.abs(r - x).max() < 1.0 and abs((r - x).mean()) < 0.05


def test_layout_counters_follow_reference():
    L = R.Layout(1 << 12, [3, 2], 2)           # two devices, 3 + 2 channels, 2 repeated
    assert L.L == 1024 and L.inc == 7 * 1024 and L.repeating_start == 5 * 1024
    assert [L.channel_base(0, c) for c in range(5)] == [0, 1024, 2048, 5120, 6144]
    assert [L.channel_base(1, c) for c in range(4)] == [3072, 4096, 5120, 6144]


def test_seeded_sampler_is_reproducible_across_processes():
    """one process per GPU: every rank constructs its own Csprng and the repeated channels (secret key, shared randomness)
    must agree, so a caller-supplied seed alone has to fix the whole (key, nonce) pair -- a random nonce per rank produced
    ciphertexts that decrypt only while device 0's limbs alone are read (found by bench.py's depth-10 circuit at N = 2)."""
    from liberate_b200.csprng import Csprng
    mk = lambda **kw: Csprng(4096, [3, 2], 2, devices=["cuda:0", "cuda:1"], local_ids=[], **kw)   # no device state: CPU-safe
    a, b = mk(seed=5), mk(seed=5)
    assert a.key == b.key and a.nonce == b.nonce == [0, 0]
    c = mk(seed=5, nonce=7)
    assert c.key == a.key and c.nonce == [7, 0]
    d, e = mk(), mk()
    assert d.key != e.key                       # os.urandom, as the reference (csprng.py:215-223)
    a.refresh()
    assert a.key != b.key
    a.refresh(5)
    assert a.key == b.key and a.nonce == b.nonce
