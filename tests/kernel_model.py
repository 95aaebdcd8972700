"""
kernel_model.py -- numpy emulation of the index algebra of liberate-fhe_b200/csrc/ntt_kernels.cuh.

Not an oracle: it mirrors OUR kernels (thread -> element mapping, round structure, twiddle index
formula) line by line so that the mapping can be checked on CPU against the oracle NTT before any
GPU time is spent.  The arithmetic itself comes from oracle.np_mont_mult.
"""
import numpy as np

from oracle import oracle as O

TILE = 4096
T = 256


def zbase(tau, p):
    return ((tau >> p) << (p + 4)) | (tau & ((1 << p) - 1))


class Limb:
    def __init__(self, P, i):
        self.q2 = np.int64(P._2q[i])
        self.m = tuple(np.int64(v[i]) for v in P.mont)

    def mont(self, a, w):
        return O.np_mont_mult(a, w, *self.m)


def ct(U, Ov, w, L):
    V = L.mont(Ov, w)
    with np.errstate(over="ignore"):
        up = U + V
        um = U + L.q2 - V
    return np.where(up < L.q2, up, up - L.q2), np.where(um < L.q2, um, um - L.q2)


def gs(U, V, w, L):
    with np.errstate(over="ignore"):
        um = U + L.q2 - V
        up = U + V
    Ov = np.where(um < L.q2, um, um - L.q2)
    return np.where(up < L.q2, up, up - L.q2), L.mont(Ov, w)


def fwd_round(e, W, s0, pre, L, first=0):
    """e: [..., 16]; pre: array broadcastable to e[..., 0]"""
    for i in range(first, 4):
        d = 8 >> i
        for k in range(16):
            if not (k & d):
                w = W[(1 << (s0 + i)) + (pre << i) + (k >> (4 - i))]
                e[..., k], e[..., k + d] = ct(e[..., k].copy(), e[..., k + d].copy(), w, L)


def inv_round(e, W, s0, pre, L, nst=4):
    for i in range(nst):
        d = 1 << i
        ip = 3 - i
        for k in range(16):
            if not (k & d):
                w = W[(1 << (s0 + ip)) + (pre << ip) + (k >> (i + 1))]
                e[..., k], e[..., k + d] = gs(e[..., k].copy(), e[..., k + d].copy(), w, L)


def field_idx(p):
    tau = np.arange(T)
    k = np.arange(16)
    return zbase(tau, p)[:, None] | (k[None, :] << p)  # [256,16] local z


def fwd_colpass(row, logN, W, L, enter_rs=None):
    """row: [N] one limb, in place. mirrors ntt_fwd_colpass"""
    b = logN - 8
    tau = np.arange(T)
    for tile in range((1 << b) // 16):
        z1 = field_idx(8)  # y = tau + 256k
        g1 = ((z1 >> 4) << b) + tile * 16 + (z1 & 15)
        e = row[g1].copy()
        if enter_rs is not None:
            e = L.mont(e, np.int64(enter_rs))
        fwd_round(e, W, 0, np.zeros(T, dtype=np.int64), L)
        sm = np.zeros(TILE, dtype=np.int64)
        sm[z1] = e
        z2 = field_idx(4)
        e = sm[z2].copy()
        fwd_round(e, W, 4, (tau >> 4).astype(np.int64), L)
        g2 = ((z2 >> 4) << b) + tile * 16 + (z2 & 15)
        row[g2] = e


def fwd_blockpass(row, logN, W, L):
    B = logN - 8
    tau = np.arange(T)
    for chunk in range((1 << logN) // TILE):
        g = row[chunk * TILE:(chunk + 1) * TILE]
        P1 = B - 4
        z = field_idx(P1)
        e = g[z].copy()
        fwd_round(e, W, logN - 4 - P1, ((chunk << (8 - P1)) | (tau >> P1)).astype(np.int64), L)
        sm = np.zeros(TILE, dtype=np.int64)
        sm[z] = e
        if B >= 8:
            P2 = B - 8
            z = field_idx(P2)
            e = sm[z].copy()
            fwd_round(e, W, logN - 4 - P2, ((chunk << (8 - P2)) | (tau >> P2)).astype(np.int64), L)
            sm[z] = e
            if B == 9:
                z = field_idx(0)
                e = sm[z].copy()
                fwd_round(e, W, logN - 4, ((chunk << 8) | tau).astype(np.int64), L, first=3)
                sm[z] = e
        elif B > 4:
            z = field_idx(0)
            e = sm[z].copy()
            fwd_round(e, W, logN - 4, ((chunk << 8) | tau).astype(np.int64), L, first=8 - B)
            sm[z] = e
        g[:] = sm


def inv_blockpass(row, logN, W, L):
    B = logN - 8
    tau = np.arange(T)
    for chunk in range((1 << logN) // TILE):
        g = row[chunk * TILE:(chunk + 1) * TILE]
        sm = g.copy()
        z = field_idx(0)
        e = sm[z].copy()
        inv_round(e, W, logN - 4, ((chunk << 8) | tau).astype(np.int64), L)
        sm[z] = e
        if B > 4:
            z = field_idx(4)
            e = sm[z].copy()
            inv_round(e, W, logN - 8, ((chunk << 4) | (tau >> 4)).astype(np.int64), L, nst=4 if B >= 8 else B - 4)
            sm[z] = e
            if B == 9:
                z = field_idx(8)
                e = sm[z].copy()
                inv_round(e, W, logN - 12, np.full(T, chunk, dtype=np.int64), L, nst=1)
                sm[z] = e
        g[:] = sm


def inv_colpass(row, logN, W, L, ninv, exit_mode, P, i):
    b = logN - 8
    tau = np.arange(T)
    q = L.q2 >> 1
    for tile in range((1 << b) // 16):
        z1 = field_idx(4)
        g1 = ((z1 >> 4) << b) + tile * 16 + (z1 & 15)
        e = row[g1].copy()
        inv_round(e, W, 4, (tau >> 4).astype(np.int64), L)
        sm = np.zeros(TILE, dtype=np.int64)
        sm[z1] = e
        z2 = field_idx(8)
        e = sm[z2].copy()
        inv_round(e, W, 0, np.zeros(T, dtype=np.int64), L)
        e = L.mont(e, np.int64(ninv))
        if exit_mode >= 1:
            e = O.np_mont_redc(e, *L.m)
        if exit_mode >= 2:
            e = np.where(e < q, e, e - q)
        if exit_mode >= 3:
            e = np.where(e <= (q >> 1), e, e - q)
        g2 = ((z2 >> 4) << b) + tile * 16 + (z2 & 15)
        row[g2] = e


def model_ntt(a, P, enter=False):
    logN = P.logN
    for i in range(a.shape[0]):
        L = Limb(P, i)
        fwd_colpass(a[i], logN, P.psi[i], L, P.Rs[i] if enter else None)
        fwd_blockpass(a[i], logN, P.psi[i], L)


def model_intt(a, P, exit_mode):
    logN = P.logN
    for i in range(a.shape[0]):
        L = Limb(P, i)
        inv_blockpass(a[i], logN, P.ipsi[i], L)
        inv_colpass(a[i], logN, P.ipsi[i], L, P.Ninv[i], exit_mode, P, i)


# ---- warp-independent block passes (ntt_fast.cuh: fast_*_block_body_w) ------------------------------------------------
def block_round_fields(B, inverse):
    """low bits p of the radix-16 fields the warp-independent block passes visit, in order"""
    if not inverse:
        if B >= 8:
            return [B - 4, B - 8] + ([0] if B == 9 else [])
        return [B - 4] + ([0] if B > 4 else [])
    if B == 4:
        return [0]
    return [0, 4] + ([5] if B == 9 else [])        # B == 9: the last level is the TOP stage of field [8:5]


def warp_of_elements(p):
    """[256,16] warp-region index (512 contiguous coefficients) of every element thread tau touches in field p"""
    return field_idx(p) // 512


def inv_blockpass_w(row, logN, W, L):
    """mirrors fast_inv_block_body_w: as inv_blockpass, but for B == 9 the last level (distance 256) is taken as the top
    stage (pairs k, k+8) of field [8:5] with ONE twiddle per 512-point sub-block, which keeps the exchange inside a warp"""
    B = logN - 8
    tau = np.arange(T)
    for chunk in range((1 << logN) // TILE):
        g = row[chunk * TILE:(chunk + 1) * TILE]
        sm = g.copy()
        z = field_idx(0)
        e = sm[z].copy()
        inv_round(e, W, logN - 4, ((chunk << 8) | tau).astype(np.int64), L)
        sm[z] = e
        if B > 4:
            z = field_idx(4)
            e = sm[z].copy()
            inv_round(e, W, logN - 8, ((chunk << 4) | (tau >> 4)).astype(np.int64), L, nst=4 if B >= 8 else B - 4)
            sm[z] = e
            if B == 9:
                z = field_idx(5)
                e = sm[z].copy()
                w = W[(1 << (logN - 9)) + ((chunk << 3) | (tau >> 5))]
                for k in range(8):
                    e[..., k], e[..., k + 8] = gs(e[..., k].copy(), e[..., k + 8].copy(), w, L)
                sm[z] = e
        g[:] = sm
