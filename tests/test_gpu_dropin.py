"""The drop-in claim, proved: the UNMODIFIED reference package (oracle/_ref/site, built from /root/reference by
oracle/build_ref.py) runs with its `liberate.ntt.ntt_cuda` extension replaced by `liberate_b200.ntt.ntt_cuda`
(src/liberate/ntt/__init__.py:1; call sites nctx.py:126-130, 532-599 and engine.py:636, 689, 701) -- keygen -> encrypt ->
mult -> rotate -> decrypt through the reference's own ckks_engine.py / ntt_context.py -- and every tensor it produces
equals, bit for bit, what the stock reference (its own CUDA kernels) produces from the same inputs.

The reference's sampler cannot be seeded, so the stock engine generates keys and ciphertexts first; the swapped engine is
a SECOND import of the same package files (the first one is moved aside in sys.modules) and works on those tensors."""
import importlib
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def same(a, b):
    return all(bool((x == y).all()) for pa, pb in zip(a.data, b.data) for x, y in zip(pa, pb))


def test_unmodified_reference_engine_on_our_ntt_cuda():
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("reference package not installed under oracle/_ref/site")
    ref_fhe, cache = ref_engine.load()
    stock_ntt_cuda = sys.modules["liberate.ntt.ntt_cuda"]
    assert "liberate_b200" not in (getattr(stock_ntt_cuda, "__file__", "") or "")
    params = dict(logN=14, num_special_primes=1, scale_bits=40, num_scales=None)      # the bronze preset
    stock = ref_fhe.ckks_engine(devices=[0], cache_folder=cache, **params)
    sk = stock.create_secret_key()
    pk = stock.create_public_key(sk)
    evk = stock.create_evk(sk)
    rotk = stock.create_rotation_key(sk, 3)
    m = stock.example(-1, 1)
    ct = stock.encorypt(m, pk)

    # second import of the very same files, with our module standing in for the extension
    from liberate_b200.ntt import ntt_cuda as ours
    saved = {k: v for k, v in sys.modules.items() if k == "liberate" or k.startswith("liberate.")}
    for k in saved:
        del sys.modules[k]
    try:
        sys.modules["liberate.ntt.ntt_cuda"] = ours
        fhe2 = importlib.import_module("liberate.fhe")
        assert sys.modules["liberate.ntt"].ntt_cuda is ours
        assert sys.modules["liberate.fhe.ckks_engine"].ntt_cuda is ours
        assert fhe2.ckks_engine is not ref_fhe.ckks_engine
        swapped = fhe2.ckks_engine(devices=[0], cache_folder=cache, **params)
        assert swapped.hash == stock.hash

        def both(fn, what):
            a, b = fn(stock), fn(swapped)
            for pi, (pa, pb) in enumerate(zip(a.data, b.data)):
                for di, (u, v) in enumerate(zip(pa, pb)):
                    if not torch.equal(u, v):
                        rows = (u != v).any(dim=1).nonzero().flatten().tolist()
                        raise AssertionError(f"{what}: polynomial {pi} device {di}: {int((u != v).sum())} elements differ in rows "
                                             f"{rows[:8]} of {u.size(0)}; first pair {u[u != v][0].item()} vs {v[u != v][0].item()}")
            return a

        # deterministic operators on shared inputs, at several levels
        x = ct
        for _ in range(3):
            trip = both(lambda e: e.cc_mult(x, x, evk, relin=False), f"cc_mult(relin=False) at level {x.level}")
            both(lambda e: e.relinearize(e.clone(trip), evk), f"relinearize at level {trip.level}")
            y = both(lambda e: e.cc_mult(x, x, evk), f"cc_mult at level {x.level}")
            both(lambda e: e.rotate_single(y, rotk), f"rotate_single at level {y.level}")
            both(lambda e: e.rescale(x), f"rescale at level {x.level}")
            x = y
        both(lambda e: e.cc_add(ct, ct), "cc_add")
        both(lambda e: e.level_up(ct, 2), "level_up")
        both(lambda e: e.mult_scalar(ct, 0.5), "mult_scalar")
        # (mc_mult is not compared: the reference's encode draws fresh random-rounding bits on every call)
        dec_a, dec_b = stock.decrode(x, sk), swapped.decrode(x, sk)
        assert np.array_equal(dec_a, dec_b)
        # key generation and encryption through the swapped engine (its own randomness): the results work in the stock engine
        sk2 = swapped.create_secret_key()
        pk2 = swapped.create_public_key(sk2)
        evk2 = swapped.create_evk(sk2)
        gk2 = swapped.create_galois_key(sk2)
        ct2 = swapped.encorypt(m, pk2)
        out = stock.rotate_galois(stock.cc_mult(ct2, ct2, evk2), gk2, 5)
        assert np.abs(stock.decrode(out, sk2) - np.roll(m * m, 5)).max() < 1e-6
        assert same(out, swapped.rotate_galois(swapped.cc_mult(ct2, ct2, evk2), gk2, 5))
        torch.cuda.synchronize()
    finally:
        for k in [k for k in sys.modules if k == "liberate" or k.startswith("liberate.")]:
            del sys.modules[k]
        sys.modules.update(saved)
