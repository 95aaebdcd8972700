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
