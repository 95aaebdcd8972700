"""
ckks_context -- parameter selection and per-prime scalar constants (host only).

Same constructor keywords and the same public attributes the engine / users read as the reference's
``ckks_context`` (src/liberate/fhe/context/ckks_context.py:151-341): q, num_scales, R, R_square, q_double,
q_lower_bits/q_higher_bits, k, k_lower_bits/k_higher_bits, R_inv, N_inv, generation_string, max_qbits,
total_qbits, ... -- identical values (tests/golden/context.json).  Differences:

  * no pickle cache and no "painted" tables: the reference materialises even/odd index tables and
    psi[C, logN, N/2] twiddles in Python-int loops (cctx.py:20-142, 317-341; ~700 MB per direction at
    platinum).  Our kernels need only a compact [C, N] table, which ntt_context builds on the GPU from the
    per-prime scalars computed here (psi_stage_factors), so construction takes milliseconds;
  * primes come from the data file liberate_b200/fhe/cache/primes.json (the reference's own tables).
"""
import math
import warnings

import numpy as np
import torch

from ..cache import tables
from ..presets import errors


def primitive_root_2N(q, N):
    """smallest-base element of order exactly 2N: g = x^((q-1)/2N) with g^N != 1 (same search order as
    the reference, cctx.py:20-28, so the same root -- and therefore the same twiddles -- is found)"""
    e = (q - 1) // (2 * N)
    g = None
    for x in range(2, N):
        g = pow(x, e, q)
        if pow(g, N, q) != 1:
            break
    return g


@errors.log_error
class ckks_context:
    def __init__(self, buffer_bit_length=62, scale_bits=40, logN=15, num_scales=None, num_special_primes=2,
                 sigma=3.2, uniform_ternary_secret=True, cache_folder=None, security_bits=128,
                 quantum="post_quantum", distribution="uniform", read_cache=True, save_cache=True,
                 verbose=False, is_secured=True):
        if buffer_bit_length != 62:
            raise errors.NotFindBufferBitLength(buffer_bit_length)
        self.generation_string = (f"{buffer_bit_length}_{scale_bits}_{logN}_{num_scales}_{num_special_primes}_"
                                  f"{security_bits}_{quantum}_{distribution}")
        self.is_secured = is_secured
        self.buffer_bit_length = buffer_bit_length
        self.scale_bits = scale_bits
        self.logN = logN
        self.N = 2 ** logN
        self.num_special_primes = num_special_primes
        self.cache_folder = cache_folder
        self.security_bits = security_bits
        self.quantum = quantum
        self.distribution = distribution
        self.sigma = sigma
        self.uniform_ternary_secret = uniform_ternary_secret
        self.secret_key_sampling_method = "uniform ternary" if uniform_ternary_secret else "sparse ternary"
        self.torch_dtype = torch.int64
        self.numpy_dtype = np.int64
        self.message_bits = buffer_bit_length - 2

        t = tables()
        try:
            message_special = t["message_special_primes"][str(self.message_bits)][str(self.N)]
        except KeyError:
            raise errors.NotFoundMessageSpecialPrimes(message_bit=self.message_bits, N=self.N)
        scale_primes = t["scale_primes"].get(f"{scale_bits},{self.N}")
        if scale_primes is None:
            raise errors.NotFoundScalePrimes(scale_bits=scale_bits, N=self.N)
        mq = t["maximum_qbits"].get(f"{security_bits},{quantum},{distribution},{logN}")
        if mq is None:
            raise errors.NotFoundScalePrimes(scale_bits=scale_bits, N=self.N)
        self.max_qbits = int(mq)

        base_special = message_special[:1 + num_special_primes]
        try:
            if num_scales is None:
                # as many scale primes as the security budget allows (cctx.py:247-257)
                budget = self.max_qbits - sum(math.log2(p) for p in base_special)
                num_scales = 0
                budget -= math.log2(scale_primes[num_scales])
                while budget > 0:
                    num_scales += 1
                    budget -= math.log2(scale_primes[num_scales])
            if num_scales > len(scale_primes):
                raise IndexError
            self.num_scales = num_scales
            self.q = list(scale_primes[:num_scales]) + list(base_special)
        except IndexError:
            raise errors.NotEnoughPrimes(scale_bits=scale_bits, N=self.N)

        self.total_qbits = math.ceil(sum(math.log2(qi) for qi in self.q))
        if self.total_qbits > self.max_qbits:
            if is_secured:
                raise errors.ViolatedAllowedQbits(scale_bits=scale_bits, N=self.N, num_scales=self.num_scales,
                                                  max_qbits=self.max_qbits, total_qbits=self.total_qbits)
            warnings.warn(f"Maximum allowed qbits are violated: max_qbits={self.max_qbits:4d} and the "
                          f"requested total is {self.total_qbits:4d}.")

        self._montgomery_constants()
        if verbose:
            self.init_print()

    def _montgomery_constants(self):
        """R = 2^62, 31-bit halves, k = (R*R^-1 - 1)/q  (cctx.py:294-315)"""
        half = self.buffer_bit_length // 2
        self.R = 2 ** self.buffer_bit_length
        self.half_buffer_bit_length = half
        self.lower_bits_mask = (1 << half) - 1
        self.full_bits_mask = self.R - 1
        self.R_square = [self.R * self.R % qi for qi in self.q]
        self.R_inv = [pow(self.R, -1, qi) for qi in self.q]
        self.k = [(self.R * ri - 1) // qi for ri, qi in zip(self.R_inv, self.q)]
        self.q_double = [2 * qi for qi in self.q]
        self.q_lower_bits = [qi & self.lower_bits_mask for qi in self.q]
        self.q_higher_bits = [qi >> half for qi in self.q]
        self.k_lower_bits = [ki & self.lower_bits_mask for ki in self.k]
        self.k_higher_bits = [ki >> half for ki in self.k]
        self.N_inv = [pow(self.N, -1, qi) for qi in self.q]

    def psi_stage_factors(self):
        """Per prime, the scalars from which the bit-reversed twiddle tables are grown by doubling:
        psi_rev[m + i] = psi_rev[i] * psi^(N/2m)  (bitrev(m+i) = bitrev(i) + N/2m for i < m = 2^s).
        Returns (fwd, inv): [C][logN] python ints, plain (not Montgomery) values of psi^(N/2^(s+1)) and
        psi^-(N/2^(s+1)).  psi is the reference's primitive 2N-th root (cctx.py:20-28, 45-54)."""
        fwd, inv = [], []
        for qi in self.q:
            g = primitive_root_2N(qi, self.N)
            gi = pow(g, -1, qi)
            fwd.append([pow(g, self.N >> (s + 1), qi) for s in range(self.logN)])
            inv.append([pow(gi, self.N >> (s + 1), qi) for s in range(self.logN)])
        return fwd, inv

    def init_print(self):
        print(f"""
ckks_context (liberate_b200):
        buffer_bit_length = {self.buffer_bit_length}   scale_bits = {self.scale_bits}   logN = {self.logN}   N = {self.N:,d}
        special primes = {self.num_special_primes}   scales = {self.num_scales}
        security = {self.security_bits} bits, {self.quantum}, {self.distribution}; secured = {self.is_secured}
        using {self.total_qbits} of at most {self.max_qbits} modulus bits
        RNS primes: {self.q}""")
