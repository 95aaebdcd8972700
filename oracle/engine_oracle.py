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


# ----------------------------------------------------------------------------------------------------
# a whole single-device engine hot path assembled from the pieces above
# ----------------------------------------------------------------------------------------------------
class OracleEngine:
    """rescale / cc_mult / relinearize / rotate on ONE device, numpy in, numpy out, following
    ckks_engine.py:746-1214 for num_devices == 1.  Rows are the device's prime order: scale primes from
    `level` on, base prime, then the K special primes (rns_partition for one device).  Used as the CPU
    baseline ("port") by bench.py and as the checker in __graft_entry__.smoke()."""

    def __init__(self, q, logN, K):
        from .oracle import Params, R
        self.q = [int(x) for x in q]
        self.logN, self.N, self.K = logN, 1 << logN, K
        self.L0 = len(self.q) - K          # ordinary limbs at level 0 (scale primes + base)
        self.S = self.L0 - 1               # number of scale primes == number of levels
        self.R = R
        self.P = Params(self.q, logN)
        nparts = -(-self.S // K)
        self.partitions = [list(range(i * K, min((i + 1) * K, self.S))) for i in range(nparts)] + [[self.S]]
        self._garner = {}

    # rows (global prime indices) alive at `level`
    def rows(self, level, special=True):
        r = list(range(level, self.L0))
        return r + list(range(self.L0, self.L0 + self.K)) if special else r

    def params(self, level, special=True):
        return self.P.slice(self.rows(level, special))

    def parts(self, level):
        """live partitions at `level` as lists of global prime indices (partially dropped ones shrink)"""
        out = []
        for p in self.partitions:
            live = [i for i in p if i >= level]
            if live:
                out.append(live)
        return out

    def rescale(self, c, level):
        """c: [L,N] rows(level, special=False) -> [L-1,N] at level+1 (engine.py:967-1052)"""
        nxt = self.params(level + 1, special=False)
        q0 = self.q[level]
        scale = np.array([pow(q0, -1, qi) * self.R % qi for qi in nxt.q], dtype=np.int64)
        return rescale_limbs(np.ascontiguousarray(c[1:]), c[0], scale, q0 // 2, nxt)

    def cc_mult(self, a, b, level):
        """a, b: (c0, c1) at `level` -> triplet (d0, d1, d2) at level+1, NTT + Montgomery (engine.py:1072-1101)"""
        lv = level + 1
        P = self.params(lv, special=False)
        xs = [self.rescale(c, level) for c in (a[0], a[1], b[0], b[1])]
        for x in xs:
            C.enter_ntt(x, P.Rs, P.psi, P._2q, *P.mont)
        return tensor_product(xs[0], xs[1], xs[2], xs[3], P), lv

    def keyswitch(self, a, ksk, level):
        """a: [L,N] plain rows(level, False); ksk: list over partitions of (k0, k1) each [E0,N] level-0 rows
        -> (c0, c1) plain [L,N]  (create_switcher, engine.py:746-904, one device)"""
        PE = self.params(level, special=True)
        E = len(PE.q)
        L = E - self.K
        acc = None
        for part in self.parts(level):
            gid = next(i for i, p in enumerate(self.partitions) if part[-1] in p)
            local = [i - level for i in part]
            m = [self.q[i] for i in part]
            key = tuple(part)
            if key not in self._garner:
                self._garner[key] = garner_constants(m, self.R)
            Y, Ls, Lc = self._garner[key]
            state = pre_extend(np.ascontiguousarray(a[local[0]:local[-1] + 1]), PE.slice(local), Y, Ls)
            L_enter = [[Lc[i] * PE.R2[t] % PE.q[t] for t in range(E)] for i in range(len(m) - 1)]
            ext = extend(state, PE, L_enter)
            C.ntt(ext, PE.psi, PE._2q, *PE.mont)
            k0, k1 = ksk[gid]
            acc = ksk_inner(ext, np.ascontiguousarray(k0[level:]), np.ascontiguousarray(k1[level:]), acc, PE)
        outs = []
        specials = self.q[-self.K:][::-1]
        PiR = [[pow(specials[i], -1, PE.q[j]) * self.R % PE.q[j] for j in range(E - i - 1)] for i in range(self.K)]
        for d in acc:
            C.intt(d, PE.ipsi, PE.Ninv, PE._2q, *PE.mont, exit_mode=2)
            outs.append(moddown(d, L, self.K, PiR, PE))
        return outs

    def relinearize(self, triplet, evk, level):
        """(d0,d1,d2) NTT+Montgomery at `level` -> (c0,c1) plain (engine.py:1117-1151)"""
        P = self.params(level, special=False)
        d = [x.copy() for x in triplet]
        for x in d:
            C.intt(x, P.ipsi, P.Ninv, P._2q, *P.mont, exit_mode=2)
        k0, k1 = self.keyswitch(d[2], evk, level)
        q = P.qa[:, None]
        out = []
        for base, ks in ((d[0], k0), (d[1], k1)):
            s = base + ks
            out.append(np.where(s < q, s, s - q))
        return out

    def mult(self, a, b, evk, level):
        t, lv = self.cc_mult(a, b, level)
        return self.relinearize(t, evk, lv), lv

    def rotate(self, ct, rotk, delta, level):
        """rotate_single (engine.py:1180-1214) for a plain ciphertext"""
        P = self.params(level, special=False)
        g = pow(3, delta % self.N, 2 * self.N)
        q = P.qa[:, None]
        moved = []
        for c in ct:
            r = automorphism(c, g, self.N) + q
            moved.append(np.where(r < q, r, r - q))
        k0, k1 = self.keyswitch(moved[1], rotk, level)
        s = C.mont_add(np.ascontiguousarray(moved[0]), k0, P._2q)
        C.reduce_2q(s, P._2q)
        return [s, k1]
