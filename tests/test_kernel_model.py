"""The thread->element mapping / twiddle-index algebra of csrc/ntt_kernels.cuh, emulated in numpy
(tests/kernel_model.py), equals the oracle NTT for every supported logN -- checked on CPU."""
import numpy as np
import pytest

import kernel_model as K
from conftest import primes_for
from oracle import oracle as O


@pytest.mark.parametrize("logN", [12, 13, 14, 15, 16, 17])
def test_two_pass_mapping_matches_oracle(logN):
    q = primes_for(logN, 1, 1)
    P = O.Params(q, logN)
    rng = np.random.default_rng(logN)
    a = np.stack([rng.integers(0, 2 * qi, 1 << logN, dtype=np.int64) for qi in q])
    a[:, ::7] -= np.array(q)[:, None] // 3
    ref = a.copy()
    O.C.ntt(ref, P.psi, P._2q, *P.mont)
    m = a.copy()
    K.model_ntt(m, P)
    assert (m == ref).all()
    ref2, m2 = a.copy(), a.copy()
    O.C.enter_ntt(ref2, P.Rs, P.psi, P._2q, *P.mont)
    K.model_ntt(m2, P, enter=True)
    assert (m2 == ref2).all()
    for mode in (0, 3):
        r, mm = ref.copy(), ref.copy()
        O.C.intt(r, P.ipsi, P.Ninv, P._2q, *P.mont, exit_mode=mode)
        K.model_intt(mm, P, mode)
        assert (mm == r).all()


@pytest.mark.parametrize("B", [4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("inverse", [False, True])
def test_warp_independent_block_passes_stay_inside_a_warp(B, inverse):
    """the invariant behind __syncwarp() in fast_*_block_body_w: in every round a thread only touches coefficients of
    the 512-coefficient slice its own warp owns (thread tau -> warp tau // 32)"""
    own = (np.arange(K.T) // 32)[:, None]
    for p in K.block_round_fields(B, inverse):
        assert (K.warp_of_elements(p) == own).all(), (B, inverse, p)


def test_inverse_last_level_as_top_stage_of_field_8_5_matches_oracle():
    """logN = 17: the warp-private form of the last inverse level (ntt_fast.cuh, B == 9) equals the oracle"""
    logN = 17
    q = primes_for(logN, 1, 0)
    P = O.Params(q, logN)
    rng = np.random.default_rng(5)
    a = np.stack([rng.integers(0, 2 * qi, 1 << logN, dtype=np.int64) for qi in q])
    ref = a.copy()
    O.C.ntt(ref, P.psi, P._2q, *P.mont)
    want = ref.copy()
    O.C.intt(want, P.ipsi, P.Ninv, P._2q, *P.mont, exit_mode=0)
    got = ref.copy()
    for i in range(got.shape[0]):
        L = K.Limb(P, i)
        K.inv_blockpass_w(got[i], logN, P.ipsi[i], L)
        K.inv_colpass(got[i], logN, P.ipsi[i], L, P.Ninv[i], 0, P, i)
    assert (got == want).all()
