"""
engine_oracle.py -- CPU restatement of the reference's hot-path ORCHESTRATION for one device:
rescale, tensor product, Garner ModUp (pre_extend / extend), evaluation-key inner product, ModDown,
Galois automorphism.  Each function follows the reference's Python sequence of ntt_cuda calls and
torch elementwise ops (src/liberate/fhe/ckks_engine.py, cited per function) on numpy int64 arrays,
using the C restatement of the kernels (oracle.C).

TEST INFRASTRUCTURE ONLY (see oracle.py).  Pinned by tests/golden/engine_D*.json: the same
functions, assembled into cc_mult / relinearize / rotate, must reproduce the tensors the unmodified
reference engine produced (tests/test_oracle_engine.py).
"""
import numpy as np

from .oracle import C


def rescale_limbs(x, r0, scale, round_at, P):
    """engine.py:1026-1038.  x: [C,N] surviving limbs; r0: [N] the dropped limb; P: Params of the C limbs."""
    with np.errstate(over="ignore"):
        d = x - r0[None, :]
    d = np.ascontiguousarray(d)
    C.mont_enter(d, np.asarray(scale, dtype=np.int64), *P.mont)
    if round_at is not None:
        d = d + (r0 > round_at).astype(np.int64)[None, :]
    d = np.ascontiguousarray(d)
    C.reduce_2q(d, P._2q)
    return d


def tensor_product(x0, x1, y0, y1, P):
    """engine.py:1095-1101"""
    d0 = C.mont_mult(x0, y0, *P.mont)
    x0y1 = C.mont_mult(x0, y1, *P.mont)
    x1y0 = C.mont_mult(x1, y0, *P.mont)
    d1 = C.mont_add(x0y1, x1y0, P._2q)
    d2 = C.mont_mult(x1, y1, *P.mont)
    return d0, d1, d2


def garner_constants(m, R):
    """ntt_context.py:323-349 for one partition with moduli m: (Y_scalar[alpha-1], L_scalar[i][j-(i+2)], L[i])"""
    alpha = len(m)
    L = [m[0]]
    for i in range(1, alpha - 1):
        L.append(L[-1] * m[i])
    Y, Ls = [], []
    for i in range(alpha - 1):
        Y.append(pow(L[i], -1, m[i + 1]) * R % m[i + 1])
        if i + 2 < alpha:
            Ls.append([(L[i] * R) % m[j] for j in range(i + 2, alpha)])
    return Y, Ls, L


def pre_extend(a_part, Ppart, Y_scalar, L_scalar):
    """engine.py:654-705 (after the optional intt).  a_part: [alpha,N]; Ppart: Params of those alpha limbs."""
    alpha = a_part.shape[0]
    state = np.repeat(a_part[0][None, :], alpha, axis=0).copy()
    for i in range(alpha - 1):
        one = Ppart.slice([i + 1])
        with np.errstate(over="ignore"):
            Y = (a_part[i + 1] - state[i + 1])[None, :].copy()
        C.mont_enter(Y, np.array([Y_scalar[i]], dtype=np.int64), *one.mont)
        state[i + 1] = Y[0]
        if i + 2 < alpha:
            rest = Ppart.slice(list(range(i + 2, alpha)))
            new_state = np.repeat(Y, alpha - (i + 2), axis=0).copy()
            C.mont_enter(new_state, np.array(L_scalar[i], dtype=np.int64), *rest.mont)
            with np.errstate(over="ignore"):
                state[i + 2:] += new_state
    return state


def extend(state, Ptarget, L_enter):
    """engine.py:707-743.  state: [alpha,N]; Ptarget: Params of the E target limbs;
    L_enter[i][t] = (L[i] * R^2) mod q_t (ntt_context.py:351-366)."""
    alpha = state.shape[0]
    E = len(Ptarget.q)
    ext = np.repeat(state[0][None, :], E, axis=0).copy()
    C.mont_enter(ext, Ptarget.Rs, *Ptarget.mont)
    for i in range(alpha - 1):
        Y = np.repeat(state[i + 1][None, :], E, axis=0).copy()
        C.mont_enter(Y, np.array(L_enter[i], dtype=np.int64), *Ptarget.mont)
        ext = C.mont_add(ext, Y, Ptarget._2q)
    return ext


def ksk_inner(ext_ntt, ksk0, ksk1, acc, Ptarget):
    """engine.py:930-937 (two mont_mult) + :832-840 (running mont_add over parts)"""
    d0 = C.mont_mult(ext_ntt, ksk0, *Ptarget.mont)
    d1 = C.mont_mult(ext_ntt, ksk1, *Ptarget.mont)
    if acc is None:
        return d0, d1
    return C.mont_add(acc[0], d0, Ptarget._2q), C.mont_add(acc[1], d1, Ptarget._2q)


def moddown(d, L, K, PiR, P):
    """engine.py:851-901.  d: [E,N] plain in [0,q) (after intt_exit_reduce); P: Params of the E rows
    (ordinary then special); PiR[i]: list of length E-i-1.  Returns the L ordinary rows.

    Faithful to the reference's out-of-bounds parameter reads: reduce_2q(d, level, -1) is launched
    over all E rows with the L-entry _2q view, whose storage continues with the special primes'
    2q -- so ALL rows are reduced with their own q (kern.cu:1062-1075 takes C from a.size(0));
    mont_enter_scalar(d, PiRi) reads garbage for the rows already consumed, which are never used."""
    E = L + K
    d = d.copy()
    ordp = P.slice(list(range(L)))
    c = d[:L]
    C.mont_enter(c, ordp.Rs, *ordp.mont)
    for i in range(K):
        Pt = np.repeat(d[E - 1 - i][None, :], E, axis=0).copy()
        C.mont_enter(Pt[:L], ordp.Rs, *ordp.mont)
        d = C.mont_sub(d, Pt, P._2q)
        live = E - i - 1
        lp = P.slice(list(range(live)))
        C.mont_enter(d[:live], np.array(PiR[i][:live], dtype=np.int64), *lp.mont)
        C.reduce_2q(d[:live], lp._2q)
    c = d[:L]
    C.mont_redc(c, *ordp.mont)
    C.reduce_2q(c, ordp._2q)
    return c.copy()


def automorphism(x, g, N):
    """encdec.py:224-246 (rotate) / :249-270 (conjugate): out[(g*j) mod N] = +-x[j]"""
    j = np.arange(N, dtype=np.int64)
    pj = (g * j) % (2 * N)
    out = np.zeros_like(x)
    sign = np.where(pj >= N, -1, 1).astype(np.int64)
    out[:, pj % N] = x * sign[None, :]
    return out
