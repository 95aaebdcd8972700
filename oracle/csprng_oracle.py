"""
csprng_oracle.py -- CPU restatement (numpy / Python integers) of the reference's sampler kernels.  TEST INFRASTRUCTURE
ONLY: nothing under liberate-fhe_b200/ imports it.

Follows, function by function:
  chacha20_block     src/liberate/csprng/chacha20_cuda_kernel.h:1-34, chacha20_cuda_kernel.cu:10-45 (20-round block
                     function on the 16-word state, feed-forward add, 64-bit counter in words 12-13)
  randint            src/liberate/csprng/randint_cuda_kernel.cu:60-101 (floor(X q / 2^128) on the 128-bit draw
                     X = (w[i+2] w[i+3] w[i] w[i+1]), four draws per block)
  discrete_gaussian  src/liberate/csprng/discrete_gaussian_cuda_kernel.cu:63-107 (sign = LSB of the high word,
                     constant-depth walk of the CDT search tree)
  randround          src/liberate/csprng/randround_cuda_kernel.cu:8-37
  build_cdt_tree     src/liberate/csprng/discrete_gaussian_sampler.py:12-116
  counters           src/liberate/csprng/csprng.py:97-108, 150-186 (per-device channel ranges, repeated channels after
                     all the others, every draw of a block advances ITS counter by `inc`)

Pinned by: the RFC 8439 section 2.3.2 block-function vector, and by tests/golden/csprng.npz -- outputs of the
reference's own chacha20_naive.chacha20 and build_CDT_binary_search_tree imported from /root/reference in the build
container (tests/golden/make_golden_csprng.py).
"""
import math

import numpy as np

CONSTANTS = (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)   # "expand 32-byte k", csprng.py:110-124
M32 = np.uint64(0xFFFFFFFF)


def _rotl(x, s):
    return ((x << np.uint64(s)) | (x >> np.uint64(32 - s))) & M32


def _qr(x, a, b, c, d):
    x[a] = (x[a] + x[b]) & M32; x[d] = _rotl(x[d] ^ x[a], 16)
    x[c] = (x[c] + x[d]) & M32; x[b] = _rotl(x[b] ^ x[c], 12)
    x[a] = (x[a] + x[b]) & M32; x[d] = _rotl(x[d] ^ x[a], 8)
    x[c] = (x[c] + x[d]) & M32; x[b] = _rotl(x[b] ^ x[c], 7)


def chacha20_states(states):
    """states: uint64 array [16, n] holding 32-bit words -> the block function's output words [16, n]"""
    s = np.asarray(states, dtype=np.uint64) & M32
    x = s.copy()
    for _ in range(10):
        _qr(x, 0, 4, 8, 12); _qr(x, 1, 5, 9, 13); _qr(x, 2, 6, 10, 14); _qr(x, 3, 7, 11, 15)
        _qr(x, 0, 5, 10, 15); _qr(x, 1, 6, 11, 12); _qr(x, 2, 7, 8, 13); _qr(x, 3, 4, 9, 14)
    return (x + s) & M32


def chacha20_block(key_nonce, counters):
    """key_nonce: 10 words (8 key + 2 nonce); counters: iterable of 64-bit block counters -> uint64 [n, 16]"""
    ctr = np.asarray(list(counters), dtype=np.uint64)
    n = len(ctr)
    st = np.zeros((16, n), dtype=np.uint64)
    for i, c in enumerate(CONSTANTS):
        st[i] = c
    for i in range(8):
        st[4 + i] = int(key_nonce[i])
    st[12] = ctr & M32
    st[13] = ctr >> np.uint64(32)
    st[14] = int(key_nonce[8])
    st[15] = int(key_nonce[9])
    return chacha20_states(st).T.copy()


def _draws(words):
    """[n,16] words -> per block 4 (low, high) pairs of Python ints"""
    out = []
    for w in words.tolist():
        out.append([((w[i] << 32) | w[i + 1], (w[i + 2] << 32) | w[i + 3]) for i in (0, 4, 8, 12)])
    return out


def randint(words, q, shift=0):
    """words [L,16] of one channel -> int64 [4L]"""
    r = [(((hi << 64) | lo) * int(q) >> 128) + shift for blk in _draws(words) for lo, hi in blk]
    return np.array(r, dtype=np.int64)


def discrete_gaussian(words, tree, size, depth):
    """tree: uint64 [2*size] (low words then high words)"""
    t = [int(v) for v in tree]
    r = []
    for blk in _draws(words):
        for lo, hi in blk:
            sign = hi & 1
            hi >>= 1
            jump, cur, cnt = 1, 0, 0
            for _ in range(depth):
                th, tl = t[cnt + cur + size], t[cnt + cur]
                ge = 1 if (hi > th or (hi == th and lo >= tl)) else 0
                cur = 2 * cur + ge
                cnt += jump
                jump *= 2
            r.append((2 * sign - 1) * cur)
    return np.array(r, dtype=np.int64)


def randround(coef, words):
    """coef float64 [n]; words [ceil(n/16),16] -> int64 [n]"""
    u = words.reshape(-1)[:len(coef)].astype(np.int64)
    a = np.abs(coef)
    fl = np.floor(a)
    ifrac = np.rint((a - fl) * 4294967296.0).astype(np.int64)
    r = fl.astype(np.int64) + (u < ifrac)
    return np.where(np.signbit(coef), -r, r).astype(np.int64)


def build_cdt_tree(security_bits=128, sigma=3.2):
    """-> (uint64 [2*size] low words | high words, size, depth)"""
    import mpmath as mpm
    mpm.mp.prec = security_bits * 2
    power = math.ceil(math.log2(6 * sigma))
    npts = 2 ** power
    s = mpm.mpf(str(sigma))
    two = mpm.mpf("2")
    S = s * mpm.sqrt(two * mpm.pi)
    prob = [mpm.exp(-mpm.mpf(str(x)) ** 2 / (two * s ** 2)) / S for x in range(npts)]
    prob[0] /= 2
    cdt = [0]
    for p in prob:
        cdt.append(cdt[-1] + p)
    cdt = [int(x * two ** mpm.mpf(str(security_bits))) for x in cdt]
    nodes = []
    for depth in range(power):
        n = 2 ** depth
        nodes += list(range(npts // n // 2, npts, npts // n))
    mask = (1 << 64) - 1
    low = [cdt[i] & mask for i in nodes]
    high = [(cdt[i] >> 64) & mask for i in nodes]
    return np.array(low + high, dtype=np.uint64), len(nodes), power


class Layout:
    """counter bookkeeping of csprng.py:97-108, 150-162: device d owns `shares[d]` channels of L blocks, then the
    `repeats` repeated channels shared by every device"""

    def __init__(self, num_coefs, shares, num_repeating):
        self.L = num_coefs // 4
        self.shares = list(shares)
        self.rep = num_repeating
        total = sum(self.shares)
        self.start = [0]
        for s in self.shares[:-1]:
            self.start.append(self.start[-1] + s * self.L)
        self.inc = (total + num_repeating) * self.L
        self.repeating_start = total * self.L

    def channel_base(self, dev, ch):
        """counter of block 0 of channel `ch` (0.. shares[dev]+rep-1) of device `dev` before any draw"""
        if ch < self.shares[dev]:
            return self.start[dev] + ch * self.L
        return self.repeating_start + (ch - self.shares[dev]) * self.L
