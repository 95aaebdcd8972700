"""GPU parity of the ENGINE: liberate_b200.fhe.ckks_engine, driven by the same seeded sampler and the same
scripted call sequence (tests/flows.py) as the unmodified reference engine was when the golden digests were
recorded (tests/golden/make_golden.py), must reproduce every key and ciphertext tensor BIT FOR BIT:
sk, pk, evk, rotation / conjugation keys, encrypt, rescale, cc_mult (triplet incl. lazy representatives),
relinearize, rotate, conjugate, add/sub, level_up, scalar ops, decrypt -- at every level, for 1, 2 and 3
logical devices (the reference's limb partitioning; several logical devices may share one physical GPU)."""
import json

import numpy as np
import pytest
import torch

import flows
from conftest import GOLDEN
from golden_utils import Checker
from seeded_rng import SeededCsprng

pytestmark = pytest.mark.gpu


def make_engine(D, params, mode="executor"):
    from liberate_b200 import fhe
    eng = fhe.ckks_engine(devices=["cuda:0"] * D, fast=(mode != "faithful"), **params)
    eng.use_executor = (mode == "executor")
    eng.rng = SeededCsprng(eng.ctx.N, [len(d) for d in eng.ntt.p.d], max(eng.ntt.num_special_primes, 2),
                           devices=eng.ntt.devices)
    return eng


@pytest.mark.parametrize("mode", ["executor", "fast-python", "faithful"])
@pytest.mark.parametrize("D", [1, 2, 3])
def test_engine_reproduces_reference_tensors(D, mode):
    """executor: mult+relin / rotate through the C executor (canonical-output FP64 + Shoup transforms, batched key
    switch, 2 C calls per mult); fast-python: the same kernels orchestrated from Python; faithful: the
    lazy-representative-faithful kernels.  All three must hit the same digests."""
    g = json.loads((GOLDEN / f"engine_D{D}.json").read_text())
    full = np.load(GOLDEN / f"engine_D{D}_full.npz")
    eng = make_engine(D, g["params"], mode)
    assert [int(x) for x in eng.ctx.q] == g["q"]
    chk = Checker(g["digests"], full, eng.ntt.devices)
    objs = flows.hot_path_flow(eng, chk)
    flows.extra_flow(eng, chk, objs)       # Galois key + rotate_galois, mc_/cm_ mult / add / sub
    assert set(chk.seen) == set(g["digests"]), set(g["digests"]) ^ set(chk.seen)
    assert not chk.failures, chk.failures[:8]
    # float side: decode of the (bit-identical) ciphertexts agrees with the reference's decode
    dec = eng.decrode(objs["ct_ab"], objs["sk"])
    assert np.abs(dec - full["decode_ab"]).max() < 1e-9
    assert np.abs(dec - full["ma"] * full["mb"]).max() < 1e-6


@pytest.mark.parametrize("mode", ["executor", "faithful"])
@pytest.mark.parametrize("D", [1, 2])
@pytest.mark.parametrize("tag", ["_sb30", "_sb42", "_sb45"])
def test_engine_reproduces_reference_tensors_at_other_scale_bits(tag, D, mode):
    """the reference's own engine tests sweep scale_bits 20..45 (src/liberate/fhe/tests/test_generate_engine.py:40-46).
    scale_bits = 30: FP64 butterflies with 2^30 primes; scale_bits = 45: every prime is >= 2^42, so the whole fused path
    (transforms, ModUp extension of multi-limb partitions of wide primes, inner product, ModDown) runs on the 64-bit
    integer kernels -- the other side of the q < 2^42 boundary of the FP64 path; scale_bits = 42: the cached primes
    alternate around 2^42, FP64 limbs and wide limbs share partitions and the key switch takes the integer pipeline."""
    g = json.loads((GOLDEN / f"engine_D{D}{tag}.json").read_text())
    full = np.load(GOLDEN / f"engine_D{D}{tag}_full.npz")
    eng = make_engine(D, g["params"], mode)
    assert [int(x) for x in eng.ctx.q] == g["q"]
    sb = g["params"]["scale_bits"]
    n_wide = sum(q >= (1 << 42) for q in eng.ctx.q[:eng.ctx.num_scales])
    assert n_wide == (0 if sb < 42 else eng.ctx.num_scales if sb > 42 else n_wide) and (sb != 42 or 0 < n_wide < eng.ctx.num_scales)
    chk = Checker(g["digests"], full, eng.ntt.devices)
    objs = flows.hot_path_flow(eng, chk)
    flows.extra_flow(eng, chk, objs)
    assert set(chk.seen) == set(g["digests"]), set(g["digests"]) ^ set(chk.seen)
    assert not chk.failures, chk.failures[:8]
    dec = eng.decrode(objs["ct_ab"], objs["sk"])
    assert np.abs(dec - full["decode_ab"]).max() < 1e-9


@pytest.mark.parametrize("D", [1, 2, 3])
def test_engine_constants_match_reference(D):
    g = json.loads((GOLDEN / f"engine_D{D}.json").read_text())
    c = json.loads((GOLDEN / f"ntt_consts_D{D}.json").read_text())
    eng = make_engine(D, g["params"])
    assert eng.hash == c["hash"]
    assert eng.ntt.starts == c["starts"] and eng.ntt.stops == c["stops"]
    assert eng.parts_alloc == c["parts_alloc"] and eng.stor_ids == c["stor_ids"]
    assert np.allclose(eng.deviations, c["deviations"], rtol=0, atol=0)
    assert np.allclose(eng.corrections, c["corrections"], rtol=0, atol=0)
    assert [t.tolist() for t in eng.mont_PR] == c["mont_PR"]
    assert [t.tolist() for t in eng.final_scalar] == c["final_scalar"]
    assert [[t.tolist() for t in lvl] for lvl in eng.rescale_scales] == c["rescale_scales"]
    assert [[[t.tolist() for t in pind] for pind in lvl] for lvl in eng.PiRs] == c["PiRs"]
    for key, ent in c["parts"].items():
        dev, rows = key.split(":")
        item = eng.ntt.parts_pack[int(dev)][tuple(int(r) for r in rows.split(","))]
        assert (item["Y_host"] or None) == ent["Y_scalar"], key
        assert (item["L_host"] or None) == ent["L_scalar"], key
        if ent["Y_scalar"] is not None:
            assert item["L_enter_host"] == ent["L_enter"], key


def test_silver_end_to_end_accuracy():
    """silver preset (BASELINE.json configs[1]): encrypt -> mult+relin -> rotate -> decrypt accuracy inside the
    reference's published envelope (cc_mult ~5e-8, rotate ~6e-9; examples/[Example] Evaluators.ipynb)"""
    from liberate_b200 import fhe
    eng = fhe.ckks_engine(**{**fhe.params["silver"], "devices": [0]})
    sk = eng.create_secret_key()
    pk = eng.create_public_key(sk)
    evk = eng.create_evk(sk)
    rotk = eng.create_rotation_key(sk, 1)
    rng = np.random.default_rng(0)
    ma = rng.uniform(-1, 1, eng.num_slots) + 1j * rng.uniform(-1, 1, eng.num_slots)
    mb = rng.uniform(-1, 1, eng.num_slots) + 1j * rng.uniform(-1, 1, eng.num_slots)
    ca, cb = eng.encorypt(ma, pk), eng.encorypt(mb, pk)
    assert np.abs(eng.decrode(ca, sk) - ma).max() < 1e-8
    prod = eng.mult(ca, cb, evk)
    assert prod.level == 1
    assert np.abs(eng.decrode(prod, sk) - ma * mb).max() < 5e-7
    rot = eng.rotate_single(prod, rotk)
    assert np.abs(eng.decrode(rot, sk) - np.roll(ma * mb, 1)).max() < 5e-7
    x = prod
    while x.level < eng.num_levels - 1:
        x = eng.mult(x, x, evk) if x.level < 3 else eng.mult(x, 1.0)
    assert x.level == eng.num_levels - 1


@pytest.mark.parametrize("D", [1, 2])
def test_captured_graph_replays_the_same_bits(D):
    """engine.capture(): a CUDA-graph replay of mult / rotate_single gives the eager result bit for bit, and picks up
    operands refreshed in place."""
    g = json.loads((GOLDEN / f"engine_D{D}.json").read_text())
    eng = make_engine(D, g["params"])
    sk = eng.create_secret_key()
    pk = eng.create_public_key(sk)
    evk = eng.create_evk(sk)
    rotk = eng.create_rotation_key(sk, 1)
    rng = np.random.default_rng(5)
    m1 = rng.uniform(-1, 1, eng.num_slots) + 1j * rng.uniform(-1, 1, eng.num_slots)
    m2 = rng.uniform(-1, 1, eng.num_slots) + 1j * rng.uniform(-1, 1, eng.num_slots)
    a, b, c = eng.encorypt(m1, pk), eng.encorypt(m2, pk), eng.encorypt(m1 * 0.5, pk)
    same = lambda x, y: all(torch.equal(u, v) for p, q in zip(x.data, y.data) for u, v in zip(p, q))
    want = eng.mult(a, b, evk)
    graph = eng.capture(eng.mult, a, b, evk)
    graph.replay()
    assert same(graph.result, want), "replayed mult differs from the eager mult"
    # refresh an operand in place: the replay must see the new data
    want2 = eng.mult(c, b, evk)
    for p, q in zip(a.data, c.data):
        for u, v in zip(p, q):
            u.copy_(v)
    graph.replay()
    assert same(graph.result, want2), "replay did not pick up the refreshed operand"
    want3 = eng.rotate_single(want2, rotk)
    g2 = eng.capture(eng.rotate_single, want2, rotk)
    g2.replay()
    assert same(g2.result, want3), "replayed rotate differs from the eager rotate"
    torch.cuda.synchronize()


@pytest.mark.parametrize("D", [1, 2])
def test_captured_circuit_with_levelling_and_scalars(D):
    """a whole circuit as one CUDA graph: (add, mult by a level-0 ciphertext -> auto level_up, scalar product, rotate) x 3.  The
    per-limb scalar tables of level_up / mult_scalar are cached on the engine, so the warm-up calls of capture() leave no
    host->device upload inside the captured region; the replay gives the eager result bit for bit."""
    g = json.loads((GOLDEN / f"engine_D{D}.json").read_text())
    eng = make_engine(D, g["params"])
    sk = eng.create_secret_key()
    pk = eng.create_public_key(sk)
    evk = eng.create_evk(sk)
    rotk = eng.create_rotation_key(sk, 1)
    rng = np.random.default_rng(9)
    m = rng.uniform(-1, 1, eng.num_slots) + 1j * rng.uniform(-1, 1, eng.num_slots)
    m /= np.abs(m).max() * 1.5
    w = 0.5 * np.exp(0.3j)
    ct, cw = eng.encorypt(m, pk), eng.encorypt(np.full(eng.num_slots, w), pk)

    def circuit(x):
        for i in range(3):
            x = eng.mult(eng.add(x, x), cw, evk)
            if i == 1:
                x = eng.mult(x, 1.0)            # mult_scalar: one more level
            x = eng.rotate_single(x, rotk)
        return x
    want = circuit(ct)
    graph = eng.capture(circuit, ct)
    graph.replay()
    torch.cuda.synchronize()
    same = all(torch.equal(u, v) for p, q in zip(graph.result.data, want.data) for u, v in zip(p, q))
    assert same, "the replayed circuit differs from the eager circuit"
    assert graph.result.level == 4
    err = np.abs(eng.decrode(graph.result, sk) - np.roll(8 * m * w ** 3, 3)).max()
    assert err < 1e-4, err


@pytest.mark.parametrize("opts", [[(12, 0)], [(18, 0)], [(17, 0)], [(17, 0), (18, 0)], [(10, 1), (9, 400)], [(9, 1)], [(16, 0)],
                                  [(19, 0)], [(22, 1)], [(22, 3)]],
                         ids=["no-rescale-fusion", "natural-order", "plain-twiddles", "plain+natural", "one-stream",
                              "many-slabs", "joint-tail", "no-pdl", "single-slab", "three-slabs"])
def test_optional_fused_paths_reproduce_reference_tensors(opts):
    """every ckks_set_option variant of the executor (rescale fusion, warp-interleaved NTT-domain order with permuted key
    copies, packed twiddles, slab and side-stream settings) must hit the same golden digests as the default path"""
    from liberate_b200._lib import lib, option_defaults
    g = json.loads((GOLDEN / "engine_D2.json").read_text())
    full = np.load(GOLDEN / "engine_D2_full.npz")
    try:
        for k, v in opts:
            lib.ckks_set_option(k, v)
        eng = make_engine(2, g["params"], "executor")
        chk = Checker(g["digests"], full, eng.ntt.devices)
        flows.hot_path_flow(eng, chk)
        assert not chk.failures, chk.failures[:8]
    finally:
        for k, v in option_defaults().items():
            lib.ckks_set_option(k, v)


def test_hoisted_rotations_decrypt_like_single_rotations():
    """engine.rotate_hoisted: one ModUp (digits -> extension -> beta*E NTTs) shared by all rotations of a ciphertext; every
    rotation then takes the inner product with the NTT-domain pre-image of its key, the inverse transforms, ModDown and the
    Galois map on the result.  Valid rotations with rotate_single's accuracy (not the same bits: Garner digits do not commute
    with the Galois map)."""
    from liberate_b200 import fhe
    for D in (1, 2):
        eng = fhe.ckks_engine(devices=["cuda:0"] * D, logN=13, num_scales=6, num_special_primes=2, scale_bits=40, is_secured=False)
        sk = eng.create_secret_key()
        pk = eng.create_public_key(sk)
        evk = eng.create_evk(sk)
        m = eng.example(-1, 1)
        ct = eng.mult(eng.encorypt(m, pk), eng.encorypt(m, pk), evk)      # level 1
        deltas = [1, 2, 8, 64, 1000]
        keys = [eng.create_rotation_key(sk, d) for d in deltas]
        outs = eng.rotate_hoisted(ct, keys)
        for d, k, o in zip(deltas, keys, outs):
            want = np.roll(m * m, d)
            single = eng.decrode(eng.rotate_single(ct, k), sk)
            got = eng.decrode(o, sk)
            assert np.abs(got - want).max() < 1e-6, (D, d)
            assert np.abs(got - single).max() < 1e-6, (D, d)
            assert o.level == ct.level and not o.ntt_state
            for poly in o.data:                     # canonical rows, like every ciphertext the engine hands out
                for dev, t in enumerate(poly):
                    assert int(t.min()) >= 0 and bool((t < eng.ntt.q[dev][eng.ntt.starts[o.level][dev]:eng.ntt.starts[o.level][dev] + t.size(0), None]).all())
        # a second call reuses the cached key pre-images and gives the same bits
        again = eng.rotate_hoisted(ct, keys)
        assert all(torch.equal(x, y) for a, b in zip(outs, again) for pa, pb in zip(a.data, b.data) for x, y in zip(pa, pb))
