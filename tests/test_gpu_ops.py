"""GPU parity: every ntt_cuda operator and every fused level-2 operator of libckks_b200.so, called through
the reference-shaped Python boundary (liberate_b200.ntt.ntt_cuda / .fused -> C ABI), is BIT-EXACT against
the CPU oracle on the same seeded inputs -- lazy [0,2q) representatives and signed inputs included."""
import numpy as np
import pytest
import torch

from conftest import primes_for
from oracle import oracle as O
from oracle import engine_oracle as EO

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def packs(P):
    t = {k: T(getattr(P, k)) for k in ("_2q", "ql", "qh", "kl", "kh", "Rs", "Ninv", "psi", "ipsi")}
    t["mont"] = [t["ql"], t["qh"], t["kl"], t["kh"]]
    t["mp"] = [t["_2q"], t["ql"], t["qh"], t["kl"], t["kh"]]
    return t


def rand_lazy(P, rng, signed=True):
    q = np.array(P.q, dtype=np.int64)[:, None]
    a = rng.integers(0, 2 * q, (len(P.q), P.N), dtype=np.int64)
    if signed:
        a[:, ::5] -= q // 3
        a[:, 1::11] -= q
    return a


def eq(t, ref):
    return bool((t.cpu().numpy() == ref).all())


@pytest.fixture(scope="module")
def nc():
    from liberate_b200.ntt import ntt_cuda
    return ntt_cuda


@pytest.mark.parametrize("logN", [12, 13, 14, 15, 16, 17])
def test_ntt_family_bit_exact(nc, logN):
    P = O.Params(primes_for(logN, 2, 2), logN)
    t = packs(P)
    rng = np.random.default_rng(100 + logN)
    a = rand_lazy(P, rng)
    dummy = [torch.empty((logN, P.N // 2), dtype=torch.int32, device=DEV)]
    # forward
    x = T(a)
    nc.ntt([x], dummy, dummy, [t["psi"]], [t["_2q"]], *[[v] for v in t["mont"]])
    ref = a.copy()
    O.C.ntt(ref, P.psi, P._2q, *P.mont)
    assert eq(x, ref), "ntt"
    # enter_ntt
    x = T(a)
    nc.enter_ntt([x], [t["Rs"]], dummy, dummy, [t["psi"]], [t["_2q"]], *[[v] for v in t["mont"]])
    ref2 = a.copy()
    O.C.enter_ntt(ref2, P.Rs, P.psi, P._2q, *P.mont)
    assert eq(x, ref2), "enter_ntt"
    # inverse family on the lazy NTT-domain data
    for mode, fn in enumerate([nc.intt, nc.intt_exit, nc.intt_exit_reduce, nc.intt_exit_reduce_signed]):
        y = T(ref)
        fn([y], dummy, dummy, [t["ipsi"]], [t["Ninv"]], [t["_2q"]], *[[v] for v in t["mont"]])
        r = ref.copy()
        O.C.intt(r, P.ipsi, P.Ninv, P._2q, *P.mont, exit_mode=mode)
        assert eq(y, r), fn.__name__


def test_ntt_accepts_painted_tables_and_strided_views(nc):
    logN = 12
    P = O.Params(primes_for(logN, 3, 3), logN)
    C, N = len(P.q), P.N
    rng = np.random.default_rng(5)
    # the reference's painted layout psi[C, logN, N/2] (ckks_context.py:336-341)
    painted = np.zeros((C, logN, N // 2), dtype=np.int64)
    ipainted = np.zeros_like(painted)
    for lvl in range(logN):
        m, tt = 1 << lvl, N >> (lvl + 1)
        painted[:, lvl, :] = np.repeat(P.psi[:, m:2 * m], tt, axis=1)
        h, t2 = N >> (lvl + 1), 1 << lvl
        ipainted[:, lvl, :] = np.repeat(P.ipsi[:, h:2 * h], t2, axis=1)
    t = packs(P)
    dummy = [torch.empty((logN, N // 2), dtype=torch.int32, device=DEV)]
    a = rand_lazy(P, rng)
    big = T(np.concatenate([a[:1] * 0 + 7, a, a[:1] * 0 + 9]))
    view = big[1:-1]  # strided row view like d[:-K] / x[start:] (engine.py:857, 926-928)
    sl = slice(1, C - 1)  # parameter views as ntt_context.param_pack makes them
    nc.enter_ntt([view[sl]], [t["Rs"][sl]], dummy, dummy, [T(painted)[sl]], [t["_2q"][sl]], *[[v[sl]] for v in t["mont"]])
    ref = a.copy()
    O.C.enter_ntt(ref[sl], P.Rs[sl], P.psi[sl], P._2q[sl], *[v[sl] for v in P.mont])
    assert eq(view, ref)
    assert int(big[0, 0]) == 7 and int(big[-1, 0]) == 9
    nc.intt_exit_reduce([view[sl]], dummy, dummy, [T(ipainted)[sl]], [t["Ninv"][sl]], [t["_2q"][sl]],
                        *[[v[sl]] for v in t["mont"]])
    O.C.intt(ref[sl], P.ipsi[sl], P.Ninv[sl], P._2q[sl], *[v[sl] for v in P.mont], exit_mode=2)
    assert eq(view, ref)


@pytest.mark.parametrize("logN", [12, 15])
def test_elementwise_ops_bit_exact(nc, logN):
    P = O.Params(primes_for(logN, 2, 2), logN)
    t = packs(P)
    rng = np.random.default_rng(7)
    a, b = rand_lazy(P, rng), rand_lazy(P, rng)
    L = lambda v: [v]
    assert eq(nc.mont_mult(L(T(a)), L(T(b)), *[[v] for v in t["mont"]])[0], O.C.mont_mult(a, b, *P.mont))
    x = T(a); nc.mont_enter(L(x), L(t["Rs"]), *[[v] for v in t["mont"]])
    r = a.copy(); O.C.mont_enter(r, P.Rs, *P.mont); assert eq(x, r)
    x = T(a); nc.mont_redc(L(x), *[[v] for v in t["mont"]])
    r = a.copy(); O.C.mont_redc(r, *P.mont); assert eq(x, r)
    for op, oop in ((nc.reduce_2q, O.C.reduce_2q), (nc.make_signed, O.C.make_signed), (nc.make_unsigned, O.C.make_unsigned)):
        x = T(a); op(L(x), L(t["_2q"]))
        r = a.copy(); oop(r, P._2q); assert eq(x, r), op.__name__
    assert eq(nc.mont_add(L(T(a)), L(T(b)), L(t["_2q"]))[0], O.C.mont_add(a, b, P._2q))
    assert eq(nc.mont_sub(L(T(a)), L(T(b)), L(t["_2q"]))[0], O.C.mont_sub(a, b, P._2q))
    e = rng.integers(-20, 20, (1, P.N), dtype=np.int64)
    assert eq(nc.tile_unsigned(L(T(e)), L(t["_2q"]))[0], O.C.tile_unsigned(e, P._2q))


@pytest.mark.parametrize("logN,alpha,K", [(12, 2, 2), (13, 3, 3), (14, 4, 4), (12, 6, 6), (12, 1, 1)])
def test_fused_keyswitch_pieces_bit_exact(logN, alpha, K):
    from liberate_b200.ntt import fused
    # limbs: alpha scale-ish primes forming one partition, then more ordinary, then K special (60-bit)
    # primes: the platinum (logN=17) table is NTT-friendly for every smaller N as well (q = 1 mod 2^18)
    import json
    from conftest import GOLDEN
    ctx = [c for c in json.loads((GOLDEN / "context.json").read_text())["contexts"] if c["args"]["logN"] == 17][0]
    qall = ctx["q"]
    small, big = qall[:-7], qall[-7:]
    nsmall = max(alpha + 1, 3)
    q = small[:nsmall] + big[:1] + big[1:K + 1]       # ordinary (nsmall scale primes + base) then K special
    Lr = nsmall + 1
    E = Lr + K
    P = O.Params(q, logN)
    t = packs(P)
    rng = np.random.default_rng(11)
    qcol = np.array(q, dtype=np.int64)[:, None]
    R = O.R

    # ---- rescale (engine.py:1026-1038)
    x = rng.integers(0, qcol[1:Lr], (Lr - 1, P.N), dtype=np.int64)
    r0 = rng.integers(0, q[0], P.N, dtype=np.int64)
    scale = np.array([pow(q[0], -1, qi) * R % qi for qi in q[1:Lr]], dtype=np.int64)
    Pr = P.slice(list(range(1, Lr)))
    ref = EO.rescale_limbs(x, r0, scale, q[0] // 2, Pr)
    got = fused.rescale(T(x), T(r0), T(scale), q[0] // 2, [v[1:Lr] for v in t["mp"]])
    assert eq(got, ref), "rescale"

    # ---- tensor product (engine.py:1095-1101), lazy signed-ish NTT-domain inputs
    ins = [rand_lazy(P, rng) for _ in range(4)]
    refs = EO.tensor_product(*ins, P)
    gots = fused.tensor_product(*[T(v) for v in ins], t["mp"])
    for g, r, n in zip(gots, refs, "012"):
        assert eq(g, r), "tensor d" + n

    # ---- Garner digits of the partition made of limbs [0, alpha) (engine.py:654-705)
    m = q[:alpha]
    Y, Ls, Lc = EO.garner_constants(m, R)
    a_part = rng.integers(0, qcol[:alpha], (alpha, P.N), dtype=np.int64)
    Pp = P.slice(list(range(alpha)))
    state_ref = EO.pre_extend(a_part, Pp, Y, Ls)
    Ltri = np.zeros((max(alpha - 1, 1), alpha), dtype=np.int64)
    for i, row in enumerate(Ls):
        for jj, v in enumerate(row):
            Ltri[i, i + 2 + jj] = v
    state = fused.garner_digits(T(a_part), T(np.array(Y, dtype=np.int64)) if alpha > 1 else None,
                                T(Ltri) if alpha > 2 else None, [v[:alpha] for v in t["mont"]])
    assert eq(state, state_ref), "garner digits"

    # ---- extend to all E limbs (engine.py:707-743)
    L_enter = [[(Lc[i] * P.R2[j]) % q[j] for j in range(E)] for i in range(alpha - 1)]
    ext_ref = EO.extend(state_ref, P, L_enter)
    ext = fused.extend(state, t["Rs"], T(np.array(L_enter, dtype=np.int64).reshape(max(alpha - 1, 0), E)) if alpha > 1 else None,
                       t["mp"])
    assert eq(ext, ext_ref), "extend"

    # ---- evaluation-key inner product over 3 fake parts (engine.py:906-937, 832-840)
    acc_ref = None
    acc0 = torch.empty((E, P.N), dtype=torch.int64, device=DEV)
    acc1 = torch.empty_like(acc0)
    for part in range(3):
        e = rand_lazy(P, rng)
        k0 = rand_lazy(P, rng, signed=False)
        k1 = rand_lazy(P, rng, signed=False)
        acc_ref = EO.ksk_inner(e, k0, k1, acc_ref, P)
        fused.ksk_accumulate(T(e), T(k0), T(k1), acc0, acc1, part == 0, t["mp"])
    assert eq(acc0, acc_ref[0]) and eq(acc1, acc_ref[1]), "ksk accumulate"

    # ---- ModDown (engine.py:851-901) on plain [0,q) rows
    d = rng.integers(0, qcol, (E, P.N), dtype=np.int64)
    Psp = q[Lr:][::-1]
    PiR = [[pow(Psp[i], -1, q[j]) * R % q[j] for j in range(E - i - 1)] for i in range(K)]
    ref = EO.moddown(d, Lr, K, PiR, P)
    PiRt = np.zeros((K, E), dtype=np.int64)
    for i in range(K):
        PiRt[i, :E - i - 1] = PiR[i]
    got = fused.moddown(T(d), Lr, K, t["Rs"], T(PiRt), t["mp"])
    assert eq(got, ref), "moddown"
    addv = rng.integers(0, qcol[:Lr], (Lr, P.N), dtype=np.int64)
    got = fused.moddown(T(d), Lr, K, t["Rs"], T(PiRt), t["mp"], add=T(addv))
    s = addv + ref
    ref_add = np.where(s < qcol[:Lr], s, s - qcol[:Lr])
    assert eq(got, ref_add), "moddown + add"

    # ---- automorphism (encdec.py:224-270 [+ engine.py:1196-1200])
    for g in (3, pow(3, 5, 2 * P.N), 2 * P.N - 1):
        xin = rng.integers(0, qcol[:Lr], (Lr, P.N), dtype=np.int64)
        ref = EO.automorphism(xin, g, P.N)
        got = fused.automorphism(T(xin), g, False)
        assert eq(got, ref), "automorphism"
        ref2 = ref + qcol[:Lr]
        ref2 = np.where(ref2 < qcol[:Lr], ref2, ref2 - qcol[:Lr])
        got = fused.automorphism(T(xin), g, True, t["_2q"][:Lr])
        assert eq(got, ref2), "automorphism+canon"


# variants of the fast transforms: library knobs (ckks_set_option) and whether the caller passes the packed tables;
# every one must give the same bits
FAST_VARIANTS = {"default": ([], True), "plain-tables": ([], False), "packed-off": ([(17, 0)], True),
                 "prefetch": ([(2, 28)], True), "one-stream": ([(10, 1)], True), "small-slabs": ([(11, 1)], True)}


@pytest.mark.parametrize("logN", [12, 13, 14, 15, 16, 17])
@pytest.mark.parametrize("force_int", [False, True])
@pytest.mark.parametrize("variant", list(FAST_VARIANTS))
def test_fast_transforms_equal_reference_sequences_after_reduction(logN, force_int, variant):
    """ckks_ntt_fast == {enter_ntt; reduce_2q}, ckks_intt_fast == intt_exit_reduce[_signed] (canonical outputs):
    FP64 error-free butterflies for the scale primes, Shoup/Harvey for the 60-bit primes."""
    from liberate_b200._lib import lib, option_defaults
    opts, packed = FAST_VARIANTS[variant]
    try:
        for k, v in opts:
            assert lib.ckks_set_option(k, v) == 0
        _check_fast_transforms(logN, force_int, packed)
    finally:
        for k, v in option_defaults().items():
            lib.ckks_set_option(k, v)


def _check_fast_transforms(logN, force_int, packed=True):
    from liberate_b200.ntt import fused
    P = O.Params(primes_for(logN, 3, 2), logN)
    t = packs(P)
    C, N = len(P.q), P.N
    rng = np.random.default_rng(300 + logN)
    q = np.array(P.q, dtype=np.int64)
    R = O.R
    tf = fused.fast_tables(T(P.psi_plain), t["_2q"] // 2)
    ti = fused.fast_tables(T(P.ipsi_plain), t["_2q"] // 2)
    # packed=True: the FastTables object (block passes read the packed last-group tables); False: the two plain tables
    sh_f, dbl_f = (tf, None) if packed else tuple(tf)
    sh_i, dbl_i = (ti, None) if packed else tuple(ti)
    qd = T(q)
    qinv = fused.reciprocals(qd) if packed else None

    def shoup(s):
        out = []
        for v, m in zip(s, q):
            w = (int(v) << 64) // int(m)
            out.append(w - (1 << 64) if w >= (1 << 63) else w)   # uint64 bit pattern as int64
        return np.array(out, dtype=np.int64)

    # forward with enter (x R): input lazy [0, 2q)
    a = rng.integers(0, 2 * q[:, None], (C, N), dtype=np.int64)
    Rp = np.array([R % int(m) for m in q], dtype=np.int64)
    x = T(a)
    fused.ntt_fast(x, sh_f, dbl_f, qd, T(Rp), T(shoup(Rp)), force_int=force_int, qinv=qinv)
    ref = a.copy()
    O.C.enter_ntt(ref, P.Rs, P.psi, P._2q, *P.mont)
    O.C.reduce_2q(ref, P._2q)
    assert eq(x, ref), "ntt_fast(enter)"
    # forward without scalar
    x = T(a)
    fused.ntt_fast(x, sh_f, dbl_f, qd, force_int=force_int, qinv=qinv)
    ref2 = a.copy()
    O.C.ntt(ref2, P.psi, P._2q, *P.mont)
    O.C.reduce_2q(ref2, P._2q)
    assert eq(x, ref2), "ntt_fast"
    # inverse with the exit chain folded into one scalar: N^-1 R^-1
    lazy = a.copy()
    O.C.ntt(lazy, P.psi, P._2q, *P.mont)            # lazy NTT-domain input in [0, 2q)
    ex = np.array([pow(N, -1, int(m)) * pow(R, -1, int(m)) % int(m) for m in q], dtype=np.int64)
    for centred, mode in ((False, 2), (True, 3)):
        y = T(lazy)
        fused.intt_fast(y, sh_i, dbl_i, qd, T(ex), T(shoup(ex)), centred=centred, force_int=force_int, qinv=qinv)
        r = lazy.copy()
        O.C.intt(r, P.ipsi, P.Ninv, P._2q, *P.mont, exit_mode=mode)
        assert eq(y, r), f"intt_fast centred={centred}"
    # warp-interleaved NTT-domain order (what the executor uses between its kernels) == the permutation of the natural result
    x = T(a)
    fused.ntt_fast(x, sh_f, dbl_f, qd, force_int=force_int, qinv=qinv, perm=True)
    assert eq(fused.perm_rows(x, inverse=True), ref2), "ntt_fast(perm)"
    y = fused.perm_rows(T(lazy))
    fused.intt_fast(y, sh_i, dbl_i, qd, T(ex), T(shoup(ex)), force_int=force_int, qinv=qinv, perm=True)
    r = lazy.copy()
    O.C.intt(r, P.ipsi, P.Ninv, P._2q, *P.mont, exit_mode=2)
    assert eq(y, r), "intt_fast(perm)"
    # batched rows: 2 x C rows share the C limbs' constants (period = C)
    reps = 5
    big = T(np.concatenate([a] * reps))
    fused.ntt_fast(big, sh_f, dbl_f, qd, period=C, force_int=force_int, qinv=qinv)
    for i in range(reps):
        assert eq(big[i * C:(i + 1) * C], ref2), f"batched period, replica {i}"
    big = T(np.concatenate([lazy] * reps))
    fused.intt_fast(big, sh_i, dbl_i, qd, T(ex), T(shoup(ex)), period=C, centred=False, force_int=force_int, qinv=qinv)
    r = lazy.copy()
    O.C.intt(r, P.ipsi, P.Ninv, P._2q, *P.mont, exit_mode=2)
    for i in range(reps):
        assert eq(big[i * C:(i + 1) * C], r), f"batched inverse, replica {i}"


@pytest.mark.parametrize("logN", [12, 16])
def test_perm_rows_is_the_documented_permutation(logN):
    """ckks_perm_rows: inside every 512-coefficient tile, coefficient 16 t + k <-> ((k >> 1) * 32 + t) * 2 + (k & 1)"""
    from liberate_b200.ntt import fused
    N = 1 << logN
    x = torch.arange(3 * N, dtype=torch.int64, device="cuda").view(3, N)
    i = np.arange(N)
    tile, t, k = i >> 9, (i >> 4) & 31, i & 15
    pos = (tile << 9) + (((k >> 1) * 32 + t) << 1) + (k & 1)
    want = np.empty((3, N), dtype=np.int64)
    want[:, pos] = x.cpu().numpy()
    y = fused.perm_rows(x)
    assert eq(y, want)
    assert eq(fused.perm_rows(y, inverse=True), x.cpu().numpy())
    # strided rows in, contiguous rows out
    big = torch.arange(6 * N, dtype=torch.int64, device="cuda").view(6, N)
    assert eq(fused.perm_rows(big[1:4]), fused.perm_rows(big[1:4].clone()).cpu().numpy())
