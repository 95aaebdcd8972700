"""
encdec -- CKKS canonical-embedding encode / decode (float64 FFT on the GPU) and the Galois
automorphisms on coefficient rows.

Same functions and numerical definition as the reference's src/liberate/fhe/encdec/encdec.py
(encode :273-297, decode :300-323, rotate :224-246, conjugate :249-270), restated in closed form:

  * slot s of the message sits at odd FFT index 3^(-s) mod 2N.  The reference obtains this placement by
    conjugating a circular shift with the folded "multiply by 3" permutation through explicit cycle
    decompositions (encdec.py:10-129, 196-206); the result is idx[s] = (3^(-s) mod 2N - 1) / 2
    (verified equal for N = 2^4..2^17, tests/test_host_logic.py);
  * encode:  coeffs = Re( FFT(sym(place(m))) * exp(-i*pi*k/N) );  decode is the inverse;
  * rotate by delta / conjugate = the automorphism X -> X^g with g = 3^delta mod 2N / g = 2N-1, done
    by one CUDA kernel (ckks_automorphism) instead of a strided scatter through a transposed view.
"""
import numpy as np
import torch

from ...ntt import fused

_idx_cache = {}
_phase_cache = {}


def slot_indices(N, device):
    key = (N, str(device))
    t = _idx_cache.get(key)
    if t is None:
        inv3 = pow(3, -1, 2 * N)
        idx = np.empty(N // 2, dtype=np.int64)
        g = 1
        for s in range(N // 2):
            idx[s] = (g - 1) // 2
            g = g * inv3 % (2 * N)
        t = torch.from_numpy(idx).to(device)
        _idx_cache[key] = t
    return t


def _phase(N, device, sign):
    key = (N, str(device), sign)
    t = _phase_cache.get(key)
    if t is None:
        k = torch.arange(N, device=device, dtype=torch.float64)
        t = torch.exp(sign * 1j * torch.pi * k / N)
        _phase_cache[key] = t
    return t


def generate_twister(N, device="cuda:0"):
    return _phase(N, device, -1)


def generate_skewer(N, device="cuda:0"):
    return _phase(N, device, +1)


def encode(m, rng=None, scale=2 ** 40, deviation=1.0, device="cuda:0", norm="forward", return_without_scaling=False):
    """m: N/2 complex (or real) slots -> N real coefficients (float64), optionally scaled and randomly
    rounded to int64 by rng.randround (reference encode, encdec.py:273-297)."""
    N = len(m) * 2
    mm = torch.from_numpy(np.array(m * deviation)).to(device)
    placed = torch.zeros((N,), dtype=mm.dtype if mm.is_complex() else torch.complex128, device=device)
    placed[slot_indices(N, device)] = mm.to(placed.dtype)
    placed = placed + placed.conj().flip(0)
    coeffs = (torch.fft.fft(placed, norm=norm) * _phase(N, device, -1)).real
    if return_without_scaling:
        return coeffs
    return rng.randround(coeffs * np.float64(scale))


def decode(m, scale=2 ** 40, correction=1.0, norm="forward", return_without_scaling=False):
    """N coefficients (tensor) -> N complex values whose first N/2 entries are the slots (encdec.py:300-323)"""
    N = len(m)
    device = m.device
    rec = torch.fft.ifft(m * _phase(N, device, +1), norm=norm)
    if not return_without_scaling:
        rec = rec / scale * correction
    out = torch.zeros_like(rec)
    out[: N // 2] = rec[slot_indices(N, device)]
    return out


def _automorph(m, g):
    N = m.size(-1)
    x = m.reshape(-1, N)
    if not x.is_contiguous():
        x = x.contiguous()
    return fused.automorphism(x, g, False).view(m.shape)


def rotate(m, delta):
    """coefficient automorphism X -> X^(3^delta): slots rotate by delta (encdec.py:224-246)"""
    N = m.size(-1)
    return _automorph(m, pow(3, delta % N, 2 * N))


def conjugate(m):
    """X -> X^(2N-1): slot-wise complex conjugation (encdec.py:249-270)"""
    N = m.size(-1)
    return _automorph(m, 2 * N - 1)
