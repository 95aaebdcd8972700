"""cpu() / cuda() / save() / load() follow the reference's wire format (src/liberate/fhe/ckks_engine.py:1790-1906, 2001-2029):
every polynomial is ONE CPU tensor in absolute prime order wrapped in a one-element list, pickled under the reference's
container class path.  CPU part: the pickle names no class of this package and loads back.  GPU part: files written by the
reference engine load here (on one and on two logical devices) and compute the same bits, and the other way round."""
import io
import pickle
import sys

import numpy as np
import pytest
import torch


def test_pickle_names_the_reference_class_and_loads_back():
    import importlib
    ce = importlib.import_module("liberate_b200.fhe.ckks_engine")
    from liberate_b200.fhe.data_struct import data_struct
    inner = data_struct([[torch.arange(4)], [torch.arange(4) + 1]], True, True, True, "pk", 0, "h")
    outer = data_struct([inner, inner], True, True, True, "ksk", 0, "h")
    had = "liberate" in sys.modules
    buf = io.BytesIO()
    with ce._wire_class() as wire:
        pickle.dump(ce._as(outer, wire), buf)
    raw = buf.getvalue()
    assert b"liberate.fhe.data_struct" in raw and b"liberate_b200" not in raw
    assert ("liberate" in sys.modules) == had, "the temporary module alias leaked"
    back = ce._as(ce._WireUnpickler(io.BytesIO(raw)).load(), data_struct)
    assert type(back) is data_struct and type(back.data[1]) is data_struct
    assert back.data[1].origin == "pk" and torch.equal(back.data[1].data[1][0], torch.arange(4) + 1)
    assert tuple(back)[1:] == tuple(outer)[1:]


def same(a, b):
    return all(bool((x == y).all()) for pa, pb in zip(a.data, b.data) for x, y in zip(pa, pb))


@pytest.mark.gpu
@pytest.mark.parametrize("D_src,D_dst", [(1, 2), (2, 1), (3, 2)])
def test_round_trip_between_device_counts(tmp_path, D_src, D_dst):
    """a ciphertext and an evaluation key saved by an engine with D_src logical devices load into one with D_dst devices"""
    from liberate_b200 import fhe
    params = dict(logN=12, num_scales=5, num_special_primes=2, scale_bits=40, is_secured=False)
    src = fhe.ckks_engine(devices=["cuda:0"] * D_src, **params)
    dst = fhe.ckks_engine(devices=["cuda:0"] * D_dst, **params)
    sk = src.create_secret_key()
    pk = src.create_public_key(sk)
    evk = src.create_evk(sk)
    m = src.example(-1, 1)
    ct = src.encorypt(m, pk)
    host = src.cpu(ct)
    assert len(host.data[0]) == 1 and host.data[0][0].device.type == "cpu"
    assert host.data[0][0].shape == (len(src.ctx.q) - src.ctx.num_special_primes, src.ctx.N)
    # prime order: row i of the CPU tensor is limb i of the chain, whatever device held it
    for dev, rows in enumerate(src.ntt.p.destination_arrays[0]):
        assert torch.equal(host.data[0][0][rows], ct.data[0][dev].cpu())
    assert same(src.cuda(host), ct)
    files = {}
    for name, obj in (("ct", ct), ("evk", evk), ("sk", sk)):
        files[name] = src.save(obj, str(tmp_path / f"{name}.pkl"))
    ct2, evk2, sk2 = (dst.load(files[n]) for n in ("ct", "evk", "sk"))
    assert ct2.data[0][0].is_cuda and len(ct2.data[0]) == D_dst
    out = dst.mult(ct2, ct2, evk2)
    assert np.abs(dst.decrode(out, sk2) - m * m).max() < 1e-6
    # and the product is the same ciphertext, limb for limb, as the one the source engine computes
    want = src.cpu(src.mult(ct, ct, evk))
    assert same(dst.cpu(out), want)


@pytest.mark.gpu
def test_files_interchange_with_the_reference_engine(tmp_path):
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("reference package not installed under oracle/_ref/site")
    ref_fhe, cache = ref_engine.load()
    from liberate_b200 import fhe
    params = dict(logN=13, num_scales=6, num_special_primes=2, scale_bits=40, is_secured=False)
    ref = ref_fhe.ckks_engine(devices=[0], cache_folder=cache, **params)
    sk = ref.create_secret_key()
    pk = ref.create_public_key(sk)
    evk = ref.create_evk(sk)
    m = ref.example(-1, 1)
    ct = ref.encorypt(m, pk)
    for name, obj in (("sk", sk), ("evk", evk), ("ct", ct)):
        ref.save(obj, str(tmp_path / f"ref_{name}.pkl"))
    want = ref.cc_mult(ct, ct, evk)
    for D in (1, 2):      # the reference wrote the files with one device; we read them with one and with two
        mine = fhe.ckks_engine(devices=["cuda:0"] * D, **params)
        sk2, evk2, ct2 = (mine.load(str(tmp_path / f"ref_{n}.pkl")) for n in ("sk", "evk", "ct"))
        out = mine.mult(ct2, ct2, evk2)
        assert same(mine.cpu(out), ref.cpu(want)), f"D={D}: product of the loaded operands differs from the reference's"
        # the other direction: our files in the reference's load()
        f = mine.save(out, str(tmp_path / f"mine_out_{D}.pkl"))
        back = ref.load(f)
        assert type(back).__module__ == "liberate.fhe.data_struct"
        assert same(back, want)
        assert np.abs(ref.decrode(back, sk) - m * m).max() < 1e-6
        fk = mine.save(mine.create_rotation_key(sk2, 1), str(tmp_path / f"mine_rotk_{D}.pkl"))
        rk = ref.load(fk)
        assert np.abs(ref.decrode(ref.rotate_single(ct, rk), sk) - np.roll(m, 1)).max() < 1e-6
