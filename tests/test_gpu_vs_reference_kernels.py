"""GPU parity against the REFERENCE'S OWN CUDA KERNELS: oracle/_ref/ntt_cuda_ref.so is the reference's ntt_cuda
extension compiled from the unmodified sources in /root/reference by oracle/build_ref.py (build container) and
shipped to the GPU box.  Every one of the 15 operators is run side by side with ours on identical inputs
(lazy [0,2q) values, signed values, strided views) and must agree BIT FOR BIT.  Skipped when the .so is absent."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import primes_for
from oracle import oracle as O

pytestmark = pytest.mark.gpu
SO = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "ntt_cuda_ref.so"
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    if not SO.exists():
        pytest.skip("oracle/_ref/ntt_cuda_ref.so not built (needs /root/reference; see oracle/build_ref.py)")
    spec = importlib.util.spec_from_file_location("ntt_cuda_ref", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ours():
    from liberate_b200.ntt import ntt_cuda
    return ntt_cuda


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def painted_tables(P):
    """the reference's table layout (ckks_context.py:89-142, 336-341) from the compact tables"""
    C, N, logN = len(P.q), P.N, P.logN
    psi = np.zeros((C, logN, N // 2), dtype=np.int64)
    ipsi = np.zeros_like(psi)
    even = np.zeros((logN, N // 2), dtype=np.int32)
    odd = np.zeros_like(even)
    ieven = np.zeros_like(even)
    iodd = np.zeros_like(even)
    b = np.arange(N // 2)
    for lvl in range(logN):
        m, t = 1 << lvl, N >> (lvl + 1)
        psi[:, lvl, :] = np.repeat(P.psi[:, m:2 * m], t, axis=1)
        even[lvl] = ((b >> (logN - 1 - lvl)) << (logN - lvl)) + (b & (t - 1))
        odd[lvl] = even[lvl] + t
        h, t2 = N >> (lvl + 1), 1 << lvl
        ipsi[:, lvl, :] = np.repeat(P.ipsi[:, h:2 * h], t2, axis=1)
        ieven[lvl] = ((b >> lvl) << (lvl + 1)) + (b & (t2 - 1))
        iodd[lvl] = ieven[lvl] + t2
    return T(psi), T(ipsi), T(even), T(odd), T(ieven), T(iodd)


@pytest.mark.parametrize("logN", [12, 14, 16])
def test_all_15_operators_match_the_reference_kernels(ref, ours, logN):
    P = O.Params(primes_for(logN, 2, 2), logN)
    C, N = len(P.q), P.N
    psi, ipsi, even, odd, ieven, iodd = painted_tables(P)
    t = {k: T(getattr(P, k)) for k in ("_2q", "ql", "qh", "kl", "kh", "Rs", "Ninv")}
    mont = [[t["ql"]], [t["qh"]], [t["kl"]], [t["kh"]]]
    rng = np.random.default_rng(logN)
    q = np.array(P.q, dtype=np.int64)[:, None]

    def lazy():
        a = rng.integers(0, 2 * q, (C, N), dtype=np.int64)
        a[:, ::5] -= q // 3
        a[:, 1::11] -= q
        return a

    a, b = lazy(), lazy()
    same = lambda x, y: bool((x == y).all())

    # in-place unary / scalar operators
    for name, args in (("mont_enter", lambda: ([t["Rs"]], *mont)), ("mont_redc", lambda: tuple(mont)),
                       ("reduce_2q", lambda: ([t["_2q"]],)), ("make_signed", lambda: ([t["_2q"]],)),
                       ("make_unsigned", lambda: ([t["_2q"]],))):
        x, y = T(a), T(a)
        getattr(ref, name)([x], *args())
        getattr(ours, name)([y], *args())
        assert same(x, y), name
    # out-of-place
    assert same(ref.mont_mult([T(a)], [T(b)], *mont)[0], ours.mont_mult([T(a)], [T(b)], *mont)[0]), "mont_mult"
    assert same(ref.mont_add([T(a)], [T(b)], [t["_2q"]])[0], ours.mont_add([T(a)], [T(b)], [t["_2q"]])[0]), "mont_add"
    assert same(ref.mont_sub([T(a)], [T(b)], [t["_2q"]])[0], ours.mont_sub([T(a)], [T(b)], [t["_2q"]])[0]), "mont_sub"
    e = rng.integers(-20, 20, (1, N), dtype=np.int64)
    assert same(ref.tile_unsigned([T(e)], [t["_2q"]])[0], ours.tile_unsigned([T(e)], [t["_2q"]])[0]), "tile_unsigned"
    # forward transforms (signed lazy inputs: the reference feeds such values, engine.py:733-740, 1163-1165)
    x, y = T(a), T(a)
    ref.ntt([x], [even], [odd], [psi], [t["_2q"]], *mont)
    ours.ntt([y], [even], [odd], [psi], [t["_2q"]], *mont)
    assert same(x, y), "ntt"
    fwd = x.clone()
    x, y = T(a), T(a)
    ref.enter_ntt([x], [t["Rs"]], [even], [odd], [psi], [t["_2q"]], *mont)
    ours.enter_ntt([y], [t["Rs"]], [even], [odd], [psi], [t["_2q"]], *mont)
    assert same(x, y), "enter_ntt"
    # inverse family
    for name in ("intt", "intt_exit", "intt_exit_reduce", "intt_exit_reduce_signed"):
        x, y = fwd.clone(), fwd.clone()
        getattr(ref, name)([x], [ieven], [iodd], [ipsi], [t["Ninv"]], [t["_2q"]], *mont)
        getattr(ours, name)([y], [ieven], [iodd], [ipsi], [t["Ninv"]], [t["_2q"]], *mont)
        assert same(x, y), name
    # strided row views (d[:-K], x[start:])
    big_r, big_o = T(np.concatenate([a, b])), T(np.concatenate([a, b]))
    sl = slice(1, 1 + C)
    tt = {k: T(np.concatenate([getattr(P, k)[1:], getattr(P, k)[:1]])) for k in ("_2q", "ql", "qh", "kl", "kh")}
    order = list(range(1, C)) + [0]
    psi_v = psi[order].contiguous()
    ref.ntt([big_r[sl]], [even], [odd], [psi_v], [tt["_2q"]], [tt["ql"]], [tt["qh"]], [tt["kl"]], [tt["kh"]])
    ours.ntt([big_o[sl]], [even], [odd], [psi_v], [tt["_2q"]], [tt["ql"]], [tt["qh"]], [tt["kl"]], [tt["kh"]])
    assert same(big_r, big_o), "ntt on a strided view"
