"""
csprng -- ChaCha20-based cryptographically secure sampler with the call surface of the reference's
``liberate.csprng.Csprng`` (src/liberate/csprng/csprng.py:18-323): randbytes / randint / discrete_gaussian /
randround / refresh, on the sm_100a kernels of csrc/csprng.cuh through the C ABI (ckks_rng_*).

Same stream as the reference for the same key and nonce (tests/test_gpu_csprng.py runs the reference's own CUDA
extensions next to these kernels on identical key material): ChaCha20 blocks in counter mode, the counter ranges
of csprng.py:97-108 (device d owns ``shares[d]`` channels of L = N/4 blocks, the repeated channels -- identical on
every device -- come after all of them), four samples per block, and every draw of a block advances that block's
counter by ``inc``.  What the reference keeps as a [blocks, 16] int64 state tensor per device (128 B per block) is
here one uint32 "epoch" per block: the state of a block is a function of (key, nonce, counter) and is rebuilt in
registers.

Differences from the reference, on purpose:
  * ``refresh(seed, nonce)`` honours a caller-supplied seed (8 words) / nonce (2 words); the reference ignores both
    and always draws from os.urandom (csprng.py:215-223).  With ``seed=None`` the key comes from os.urandom as there.
  * ``local_ids``: in the one-process-per-GPU model a process only creates state for the logical devices it owns;
    entries of the returned per-device lists for other devices are None.  Every process must then be constructed
    with the SAME seed and nonce (the engine broadcasts them), or the repeated channels would differ between ranks.
"""
import ctypes
import math
import os

import numpy as np
import torch

from .._lib import check, lib

# CDT search tree of the discrete Gaussian for the default parameters (128 bits, sigma = 3.2): 31 low words, then the
# 31 high words (what discrete_gaussian_sampler.build_CDT_binary_search_tree returns, flattened).  Other sigmas are
# computed with mpmath exactly like the reference does.
_DEFAULT_TREE = None


def build_cdt_tree(security_bits=128, sigma=3.2):
    """CDT of the half Gaussian at 2^ceil(log2(6 sigma)) points in 128-bit fixed point, laid out as a binary search
    tree (discrete_gaussian_sampler.py:12-116) -> (uint64 [2*size]: low words | high words, size, depth)"""
    import mpmath as mpm
    old = mpm.mp.prec
    mpm.mp.prec = security_bits * 2
    try:
        power = math.ceil(math.log2(6 * sigma))
        points = 2 ** power
        s, two = mpm.mpf(str(sigma)), mpm.mpf("2")
        norm = s * mpm.sqrt(two * mpm.pi)
        prob = [mpm.exp(-mpm.mpf(str(x)) ** 2 / (two * s ** 2)) / norm for x in range(points)]
        prob[0] /= 2                                    # half plane: the mass at 0 is shared with the mirror image
        cdt = [0]
        for p in prob:
            cdt.append(cdt[-1] + p)
        cdt = [int(x * two ** mpm.mpf(str(security_bits))) for x in cdt]
    finally:
        mpm.mp.prec = old
    nodes = []
    for depth in range(power):
        n = 2 ** depth
        nodes += list(range(points // n // 2, points, points // n))
    mask = (1 << 64) - 1
    flat = [cdt[i] & mask for i in nodes] + [(cdt[i] >> 64) & mask for i in nodes]
    return np.array(flat, dtype=np.uint64), len(nodes), power


def _words(value, count, what):
    """seed / nonce material -> `count` 32-bit words"""
    if value is None:
        return [int.from_bytes(os.urandom(4), "big") for _ in range(count)]
    if isinstance(value, (int, np.integer)):
        v = int(value)
        return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(count)]
    vals = [int(x) & 0xFFFFFFFF for x in value]
    if len(vals) != count:
        raise ValueError(f"{what} must be an integer or {count} 32-bit words")
    return vals


class Csprng:
    def __init__(self, num_coefs=2 ** 15, num_channels=[8], num_repeating_channels=2, sigma=3.2, devices=None,
                 seed=None, nonce=None, local_ids=None):
        self.num_coefs = num_coefs
        self.num_channels = list(num_channels)
        self.num_repeating_channels = num_repeating_channels
        self.sigma = sigma
        if devices is None:
            devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
        self.devices = list(devices)
        self.num_devices = len(self.devices)
        self.local_ids = list(range(self.num_devices)) if local_ids is None else list(local_ids)
        if len(self.num_channels) == 1:
            self.shares = [self.num_channels[0]] * self.num_devices
        elif len(self.num_channels) == self.num_devices:
            self.shares = list(self.num_channels)
        else:
            raise Exception("There was a contradicting mismatch between num_channels, and devices.")
        self.total_num_channels = sum(self.shares)
        self.L = self.num_coefs // 4                    # one ChaCha20 block = four samples
        global _DEFAULT_TREE
        if sigma == 3.2:
            if _DEFAULT_TREE is None:
                _DEFAULT_TREE = build_cdt_tree(128, 3.2)
            self.btree, self.btree_size, self.tree_depth = _DEFAULT_TREE
        else:
            self.btree, self.btree_size, self.tree_depth = build_cdt_tree(128, sigma)
        self._lut = (ctypes.c_uint64 * len(self.btree))(*[int(v) for v in self.btree])
        # counter ranges (csprng.py:97-108)
        self.start_ind = [0]
        for s in self.shares[:-1]:
            self.start_ind.append(self.start_ind[-1] + s * self.L)
        self.inc = (self.total_num_channels + self.num_repeating_channels) * self.L
        self.repeating_start = self.total_num_channels * self.L
        self._ctr_base, self._epoch = {}, {}
        for d in self.local_ids:
            base = [self.start_ind[d] + c * self.L for c in range(self.shares[d])]
            base += [self.repeating_start + r * self.L for r in range(self.num_repeating_channels)]
            self._ctr_base[d] = torch.tensor(base, dtype=torch.int64, device=self.devices[d])
            self._epoch[d] = torch.zeros(len(base) * self.L, dtype=torch.int32, device=self.devices[d])
        self._q_cache = {}
        self.refresh(seed, nonce)

    # ------------------------------------------------------------------------------------------------------
    def refresh(self, seed=None, nonce=None):
        """new key / nonce (os.urandom unless given) and all block counters back to their start"""
        self.key = _words(seed, 8, "seed")
        # a caller-supplied seed without a nonce gives a reproducible stream (zero nonce), not a half-random one: one
        # process per GPU, the ranks must agree on BOTH words or their repeated channels differ
        self.nonce = _words(0 if (nonce is None and seed is not None) else nonce, 2, "nonce")
        self._kn = (ctypes.c_uint32 * 10)(*(self.key + self.nonce))
        for e in self._epoch.values():
            e.zero_()

    def _range(self, d, share, repeats):
        """channels [shares[d] - share, shares[d] + repeats) of device d (csprng.py:222-229)"""
        lo, hi = self.shares[d] - share, self.shares[d] + repeats
        if lo < 0 or repeats > self.num_repeating_channels:
            raise ValueError("more channels requested than the sampler was built for")
        return lo, hi

    def _ptrs(self, d, lo):
        return self._ctr_base[d].data_ptr() + 8 * lo, self._epoch[d].data_ptr() + 4 * lo * self.L

    def _stream(self, d):
        return torch.cuda.current_stream(torch.device(self.devices[d])).cuda_stream

    # ------------------------------------------------------------------------------------------------------
    def randbytes(self, shares=None, repeats=0, reshape=False):
        """raw ChaCha20 output: per device an int64 tensor [(share + repeats) * L, 16] of 32-bit words"""
        shares = self.shares if shares is None else shares
        out = []
        for d in range(self.num_devices):
            if d not in self.local_ids:
                out.append(None)
                continue
            lo, hi = self._range(d, shares[d], repeats)
            t = torch.empty(((hi - lo) * self.L, 16), dtype=torch.int64, device=self.devices[d])
            if hi > lo:
                cb, ep = self._ptrs(d, lo)
                with torch.cuda.device(t.device):
                    check(lib.ckks_rng_bytes(t.data_ptr(), hi - lo, self.L, self._kn, cb, ep, self.inc, self._stream(d)), "rng_bytes")
            out.append(t.view(-1, self.L, 16) if reshape else t)
        return out

    def randint(self, amax=3, shift=0, repeats=0):
        """uniform integers in [shift, amax + shift): amax is a number (one repeated channel per device by default
        usage) or, per device, the list of moduli of the channels to draw; the last `repeats` entries are drawn
        from the repeated channels and are identical on every device"""
        if not isinstance(amax, (list, tuple)):
            amax = [[amax] for _ in self.shares]
        out = []
        for d, am in enumerate(amax):
            if d not in self.local_ids:
                out.append(None)
                continue
            lo, hi = self._range(d, len(am) - repeats, repeats)
            key = (d, tuple(int(x) for x in am))
            q = self._q_cache.get(key)
            if q is None:
                q = torch.from_numpy(np.array([int(x) for x in am], dtype=np.uint64).view(np.int64)).to(self.devices[d])
                if len(self._q_cache) < 256:
                    self._q_cache[key] = q
            t = torch.empty((hi - lo, self.num_coefs), dtype=torch.int64, device=self.devices[d])
            if hi > lo:
                cb, ep = self._ptrs(d, lo)
                with torch.cuda.device(t.device):
                    check(lib.ckks_rng_randint(t.data_ptr(), hi - lo, self.L, q.data_ptr(), int(shift), self._kn, cb, ep,
                                               self.inc, self._stream(d)), "rng_randint")
            out.append(t)
        return out

    def discrete_gaussian(self, non_repeats=0, repeats=1):
        """rounded Gaussian of standard deviation sigma, truncated at 2^ceil(log2(6 sigma))"""
        shares = non_repeats if isinstance(non_repeats, (list, tuple)) else [non_repeats] * self.num_devices
        out = []
        for d in range(self.num_devices):
            if d not in self.local_ids:
                out.append(None)
                continue
            lo, hi = self._range(d, shares[d], repeats)
            t = torch.empty((hi - lo, self.num_coefs), dtype=torch.int64, device=self.devices[d])
            if hi > lo:
                cb, ep = self._ptrs(d, lo)
                with torch.cuda.device(t.device):
                    check(lib.ckks_rng_gaussian(t.data_ptr(), hi - lo, self.L, self._lut, self.btree_size, self.tree_depth,
                                                self._kn, cb, ep, self.inc, self._stream(d)), "rng_gaussian")
            out.append(t)
        return out

    def randround(self, coef):
        """sign(x) (floor|x| + Bernoulli(frac|x|)) as int64, 32 random bits per coefficient from the first blocks of
        the first local device's first channel (csprng.py:307-323); coef: float64 tensor on that device"""
        d = self.local_ids[0]
        coef = coef.contiguous()
        if coef.dtype != torch.float64 or str(coef.device) != str(torch.device(self.devices[d])):
            raise ValueError("randround expects a float64 tensor on the sampler's first local device")
        n = coef.numel()
        if (n + 15) // 16 > self.L:
            raise ValueError("randround: more coefficients than one channel holds")
        out = torch.empty(coef.shape, dtype=torch.int64, device=coef.device)
        cb, ep = self._ptrs(d, 0)
        with torch.cuda.device(coef.device):
            check(lib.ckks_rng_randround(coef.data_ptr(), out.data_ptr(), n, self._kn, cb, ep, self.inc, self._stream(d)),
                  "rng_randround")
        return out
