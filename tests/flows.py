"""
flows.py -- one scripted sequence of public ckks_engine calls, executed
  (a) by the UNMODIFIED reference engine in the build container (tests/golden/make_golden.py), and
  (b) by liberate_b200's engine in the tests,
both driven by tests/seeded_rng.SeededCsprng, so every key / ciphertext tensor can be compared
bit for bit (SURVEY.md 8c pin (3): "same sk/pk/evk/rotk/ct tensors to both engines").

``rec(name, obj)`` records (generation) or checks (test) an object; ``rec.fix(name, obj)`` does the
same but returns the golden value so that FFT-dependent inputs (encode) are identical downstream.
"""
import numpy as np


def messages(num_slots, seed=7):
    r = np.random.default_rng(seed)
    ma = r.uniform(-1, 1, num_slots) + 1j * r.uniform(-1, 1, num_slots)
    mb = r.uniform(-1, 1, num_slots) + 1j * r.uniform(-1, 1, num_slots)
    return ma, mb


def hot_path_flow(eng, rec, deep=True):
    """keys -> encrypt -> rescale / cc_mult / relinearize / rotate at several levels."""
    ma, mb = messages(eng.num_slots)
    sk = eng.create_secret_key()
    rec("sk", sk)
    pk = eng.create_public_key(sk)
    rec("pk", pk)
    evk = eng.create_evk(sk)
    rec("evk", evk)
    rotk1 = eng.create_rotation_key(sk, 1)
    rec("rotk1", rotk1)
    conjk = eng.create_conjugation_key(sk)
    rec("conjk", conjk)

    pt_a = rec.fix("pt_a", eng.encode(ma, 0))
    pt_b = rec.fix("pt_b", eng.encode(mb, 0))
    ct_a = eng.encrypt(pt_a, pk, 0)
    rec("ct_a", ct_a)
    ct_b = eng.encrypt(pt_b, pk, 0)
    rec("ct_b", ct_b)

    rec("rescale_a", eng.rescale(ct_a))
    ctt = eng.cc_mult(ct_a, ct_b, evk, relin=False)
    rec("ctt_ab", ctt)
    ct_ab = eng.relinearize(ctt, evk)
    rec("ct_ab", ct_ab)
    rec("pt_dec_ab", eng.decrypt(ct_ab, sk))
    rec("pt_dec_ctt", eng.decrypt(ctt, sk))

    rec("rot_a", eng.rotate_single(ct_a, rotk1))
    rec("rot_ab", eng.rotate_single(ct_ab, rotk1))
    rec("conj_a", eng.conjugate(ct_a, conjk))
    rec("add_ab", eng.cc_add(ct_a, ct_b))
    rec("sub_ab", eng.cc_sub(ct_a, ct_b))
    rec("add_ctt", eng.cc_add(ctt, ctt))
    rec("levelup_a_2", eng.level_up(ct_a, 2))
    rec("auto_add", eng.add(ct_a, ct_ab))
    rec("auto_mult", eng.mult(ct_a, ct_ab, evk))
    rec("negate_a", eng.negate(ct_a))
    rec("mult_int", eng.mult(ct_a, 3))
    rec("mult_float", eng.mult(ct_a, 0.5))
    rec("add_float", eng.add(ct_a, 1.25))
    rec("sub_float", eng.sub(ct_a, 0.75))

    if deep:
        x = ct_ab
        while x.level < eng.num_levels - 1:
            x = eng.cc_mult(x, x, evk)
            rec(f"square_l{x.level}", x)
            rec(f"rot_l{x.level}", eng.rotate_single(x, rotk1))
        rec("pt_dec_deep", eng.decrypt(x, sk))
    return dict(sk=sk, pk=pk, evk=evk, rotk1=rotk1, ct_a=ct_a, ct_b=ct_b, ct_ab=ct_ab, ma=ma, mb=mb)


class fixed_encode:
    """engine.encode is FFT-dependent (and random-rounds): inside mc_mult / mc_add the engine calls it itself, so for the
    duration of such a call route it through rec.fix -- recorded in generation, replaced by the golden value in tests"""

    def __init__(self, eng, rec, name):
        self.eng, self.rec, self.name = eng, rec, name

    def __enter__(self):
        orig = self.eng.encode
        self.orig = orig
        rec, name = self.rec, self.name

        def enc(m, level=0, padding=True):
            return rec.fix(name, orig(m, level, padding))
        self.eng.encode = enc
        return self

    def __exit__(self, *exc):
        del self.eng.encode          # back to the class method
        return False


def extra_flow(eng, rec, objs):
    """Galois keys and composite rotations, plaintext (message) operands: create_galois_key (engine.py:1216-1232),
    rotate_galois (:1234-1266), mc_mult / mc_add / mc_sub and their cm_ twins through the dispatchers (:2052-2219)."""
    sk, ct_a, ct_ab, mb = objs["sk"], objs["ct_a"], objs["ct_ab"], objs["mb"]
    gk = eng.create_galois_key(sk)
    rec("galk", gk)
    rec("rotg_a_5", eng.rotate_galois(ct_a, gk, 5))
    rec("rotg_ab_m3", eng.rotate_galois(ct_ab, gk, -3))
    with fixed_encode(eng, rec, "pt_mc_mult"):
        rec("mc_mult", eng.mult(mb, ct_a))
    with fixed_encode(eng, rec, "pt_cm_mult"):
        rec("cm_mult_l1", eng.mult(ct_ab, mb))
    with fixed_encode(eng, rec, "pt_mc_add"):
        rec("mc_add", eng.add(mb, ct_a))
    with fixed_encode(eng, rec, "pt_cm_sub"):
        rec("cm_sub", eng.sub(ct_a, mb))
    with fixed_encode(eng, rec, "pt_mc_sub"):
        rec("mc_sub", eng.sub(mb, ct_ab))
    return dict(galk=gk)
