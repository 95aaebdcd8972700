"""GPU parity of the sampler kernels (csrc/csprng.cuh through liberate_b200.csprng.Csprng) with the CPU oracle
(oracle/csprng_oracle.py, pinned by RFC 8439 and the reference's Python code) and -- when the reference package is
installed under oracle/_ref/site -- with the reference's own CUDA extensions on identical key material."""
import numpy as np
import pytest
import torch

from oracle import csprng_oracle as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KEY = [0x03020100, 0x07060504, 0x0B0A0908, 0x0F0E0D0C, 0x13121110, 0x17161514, 0x1B1A1918, 0x1F1E1D1C]
NONCE = [0x4A000000, 0x12345678]


def make(N=1 << 12, shares=(3, 2), rep=2, **kw):
    from liberate_b200.csprng import Csprng
    return Csprng(N, list(shares), rep, devices=[DEV] * len(shares), seed=KEY, nonce=NONCE, **kw)


def words(lay, dev, ch, epoch, L):
    base = lay.channel_base(dev, ch) + epoch * lay.inc
    return R.chacha20_block(KEY + NONCE, range(base, base + L))


def test_randbytes_counters_and_epochs():
    g = make()
    lay = R.Layout(g.num_coefs, g.shares, g.num_repeating_channels)
    for epoch in range(2):                                   # second call: every block's counter moved by inc
        got = g.randbytes(repeats=2, reshape=True)
        for dev in range(2):
            n_ch = g.shares[dev] + 2
            assert got[dev].shape == (n_ch, g.L, 16)
            for ch in range(n_ch):
                want = words(lay, dev, ch, epoch, g.L)
                assert (got[dev][ch].cpu().numpy().astype(np.uint64) == want).all(), (epoch, dev, ch)
    # repeated channels are identical on every device
    assert torch.equal(got[0][-2:], got[1][-2:])
    # partial draw: only the last channel of device 0 and one repeated channel advance
    part = g.randbytes(shares=[1, 0], repeats=1, reshape=True)
    assert (part[0][0].cpu().numpy().astype(np.uint64) == words(lay, 0, 2, 2, g.L)).all()
    assert (part[0][1].cpu().numpy().astype(np.uint64) == words(lay, 0, 3, 2, g.L)).all()
    assert (part[1][0].cpu().numpy().astype(np.uint64) == words(lay, 1, 2, 2, g.L)).all()


def test_randint_gaussian_randround_bit_exact_vs_oracle():
    g = make()
    lay = R.Layout(g.num_coefs, g.shares, g.num_repeating_channels)
    q0 = [1099511799809, 1152921504606830593, 3]
    q1 = [1152921504606830593, 3]
    out = g.randint([q0, q1], shift=-1, repeats=1)          # device 0: channels 1,2 + repeated 0; device 1: channel 1 + repeated 0
    for dev, qs, chans in ((0, q0, (1, 2, 3)), (1, q1, (1, 2))):
        for row, (q, ch) in enumerate(zip(qs, chans)):
            want = R.randint(words(lay, dev, ch, 0, g.L), q, shift=-1)
            assert (out[dev][row].cpu().numpy() == want).all(), ("randint", dev, ch)
    assert torch.equal(out[0][-1], out[1][-1])
    tree, size, depth = R.build_cdt_tree()
    assert (np.asarray(g.btree) == tree).all() and (g.btree_size, g.tree_depth) == (size, depth)
    gs = g.discrete_gaussian(non_repeats=[1, 2], repeats=2)
    # device 0 drew channel 2 (second draw of it) + repeated 0 (second draw), repeated 1 (first draw)
    assert (gs[0][0].cpu().numpy() == R.discrete_gaussian(words(lay, 0, 2, 1, g.L), tree, size, depth)).all()
    assert (gs[0][1].cpu().numpy() == R.discrete_gaussian(words(lay, 0, 3, 1, g.L), tree, size, depth)).all()
    assert (gs[0][2].cpu().numpy() == R.discrete_gaussian(words(lay, 0, 4, 0, g.L), tree, size, depth)).all()
    assert (gs[1][0].cpu().numpy() == R.discrete_gaussian(words(lay, 1, 0, 0, g.L), tree, size, depth)).all()
    assert (gs[1][1].cpu().numpy() == R.discrete_gaussian(words(lay, 1, 1, 1, g.L), tree, size, depth)).all()
    assert abs(float(gs[0].double().std()) - 3.2) < 0.1
    # randround: first N/16 blocks of device 0 / channel 0 (never drawn so far)
    rng = np.random.default_rng(1)
    x = rng.uniform(-1e6, 1e6, g.num_coefs)
    x[:4] = [0.0, -0.0, 2.5, -2.5]
    got = g.randround(torch.from_numpy(x).to(DEV)).cpu().numpy()
    want = R.randround(x, words(lay, 0, 0, 0, g.num_coefs // 16))
    assert (got == want).all()


def test_refresh_restarts_the_stream_and_honours_seed():
    g = make()
    a = g.randint(amax=3, shift=-1, repeats=1)
    g.refresh(KEY, NONCE)
    b = g.randint(amax=3, shift=-1, repeats=1)
    assert torch.equal(a[0], b[0])
    g.refresh()                                              # os.urandom key
    c = g.randint(amax=3, shift=-1, repeats=1)
    assert not torch.equal(a[0], c[0])


def test_same_stream_as_the_reference_cuda_extensions():
    """the reference's Csprng with OUR key and nonce written into its state tensors produces the same samples"""
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("reference package not installed under oracle/_ref/site")
    ref_engine.load()
    from liberate.csprng import Csprng as RefCsprng
    N, shares, rep = 1 << 12, [3, 2], 2
    ref = RefCsprng(N, shares, rep, devices=[DEV, DEV])
    ref.key = [torch.tensor(KEY, dtype=torch.int64, device=DEV) for _ in range(2)]
    ref.nonce = [torch.tensor(NONCE, dtype=torch.int64, device=DEV) for _ in range(2)]
    for d in range(2):
        ref.initialize_states(d)
    g = make(N, shares, rep)
    q0 = [1099511799809, 1152921504606830593, 1073741827]
    q1 = [1152921504606830593, 1073741827]
    for _ in range(2):
        r_ref = ref.randint([q0, q1], shift=0, repeats=1)
        r_our = g.randint([q0, q1], shift=0, repeats=1)
        for d in range(2):
            assert torch.equal(r_ref[d], r_our[d]), "randint"
        g_ref = ref.discrete_gaussian(non_repeats=[1, 2], repeats=2)
        g_our = g.discrete_gaussian(non_repeats=[1, 2], repeats=2)
        for d in range(2):
            assert torch.equal(g_ref[d], g_our[d]), "discrete_gaussian"
        b_ref = ref.randbytes(repeats=1)
        b_our = g.randbytes(repeats=1)
        for d in range(2):
            assert torch.equal(b_ref[d], b_our[d]), "randbytes"
    x = torch.from_numpy(np.random.default_rng(2).uniform(-1e5, 1e5, N)).to(DEV)
    assert torch.equal(ref.randround(x.clone()), g.randround(x)), "randround"
