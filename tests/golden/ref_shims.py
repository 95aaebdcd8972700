"""
ref_shims.py -- lets the UNMODIFIED reference Python package (/root/reference/src/liberate) be
imported and driven on CPU in the build container, so that golden vectors can be produced by the
reference's own ckks_context / rns_partition / ntt_context / ckks_engine code.

Used only by tests/golden/make_golden.py (build container; /root/reference does not exist on the
GPU box).  Nothing here is product code.

What is stubbed:
  * matplotlib            -- absent; imported at module scope by generate_primes.py / helpers.py
  * numpy.bool8           -- removed in numpy 2
  * liberate.ntt.ntt_cuda -- the 15-op CUDA extension; replaced by the CPU oracle (oracle/ckks_oracle.c),
                             which restates the same kernels, so the reference's ORCHESTRATION (Garner
                             ModUp, ModDown, rescale, tensor product, partitioning) runs unmodified
  * liberate.csprng       -- 4 CUDA extensions; replaced by a seeded numpy sampler with the same
                             method signatures/shapes (values need only be valid samples; the
                             reference's generator cannot be seeded anyway)
  * torch.Tensor.pin_memory / .cuda -- no CUDA runtime here; identity / .to(device)
"""
import sys
import types
import shutil
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[2]
REF_SRC = Path("/root/reference/src")
sys.path.insert(0, str(REPO))
from oracle import oracle as O  # noqa: E402


def _np(t):
    return t.numpy()


class _CompactCache:
    """painted psi[C, logN, N/2] (cctx.py:336-341) -> compact [C, N] bit-reversed table"""

    def __init__(self):
        self.cache = {}

    def get(self, psi, forward):
        key = (psi.data_ptr(), tuple(psi.shape), tuple(psi.stride()), forward)
        hit = self.cache.get(key)
        if hit is not None:
            return hit
        p = psi.numpy()
        C_, logN, half = p.shape
        N = half * 2
        out = np.zeros((C_, N), dtype=np.int64)
        for lvl in range(logN):
            if forward:
                m = 1 << lvl
                t = N >> (lvl + 1)
                out[:, m:2 * m] = p[:, lvl, ::t][:, :m]
            else:
                h = N >> (lvl + 1)
                t = 1 << lvl
                out[:, h:2 * h] = p[:, lvl, ::t][:, :h]
        self.cache[key] = out
        return out


_cc = _CompactCache()


def make_ntt_cuda_module():
    m = types.ModuleType("liberate.ntt.ntt_cuda")
    n = lambda ts: [t.numpy() for t in ts]

    def mont_mult(a, b, ql, qh, kl, kh):
        return [torch.from_numpy(O.C.mont_mult(_np(x), _np(y), _np(l), _np(h), _np(k), _np(kk)))
                for x, y, l, h, k, kk in zip(a, b, ql, qh, kl, kh)]

    def mont_enter(a, Rs, ql, qh, kl, kh):
        for x, r, l, h, k, kk in zip(a, Rs, ql, qh, kl, kh):
            O.C.mont_enter(_np(x), np.ascontiguousarray(_np(r)), _np(l), _np(h), _np(k), _np(kk))

    def ntt(a, even, odd, psi, _2q, ql, qh, kl, kh):
        for x, p, q2, l, h, k, kk in zip(a, psi, _2q, ql, qh, kl, kh):
            O.C.ntt(_np(x), _cc.get(p, True), _np(q2), _np(l), _np(h), _np(k), _np(kk))

    def enter_ntt(a, Rs, even, odd, psi, _2q, ql, qh, kl, kh):
        mont_enter(a, Rs, ql, qh, kl, kh)
        ntt(a, even, odd, psi, _2q, ql, qh, kl, kh)

    def _intt(mode):
        def f(a, even, odd, psi, Ninv, _2q, ql, qh, kl, kh):
            for x, p, ni, q2, l, h, k, kk in zip(a, psi, Ninv, _2q, ql, qh, kl, kh):
                O.C.intt(_np(x), _cc.get(p, False), _np(ni), _np(q2), _np(l), _np(h), _np(k), _np(kk), mode)
        return f

    def mont_redc(a, ql, qh, kl, kh):
        for x, l, h, k, kk in zip(a, ql, qh, kl, kh):
            O.C.mont_redc(_np(x), _np(l), _np(h), _np(k), _np(kk))

    def reduce_2q(a, _2q):
        for x, q2 in zip(a, _2q):
            O.C.reduce_2q(_np(x), _np(q2))

    def make_signed(a, _2q):
        for x, q2 in zip(a, _2q):
            O.C.make_signed(_np(x), _np(q2))

    def make_unsigned(a, _2q):
        for x, q2 in zip(a, _2q):
            O.C.make_unsigned(_np(x), _np(q2))

    def mont_add(a, b, _2q):
        return [torch.from_numpy(O.C.mont_add(_np(x), _np(y), _np(q2))) for x, y, q2 in zip(a, b, _2q)]

    def mont_sub(a, b, _2q):
        return [torch.from_numpy(O.C.mont_sub(_np(x), _np(y), _np(q2))) for x, y, q2 in zip(a, b, _2q)]

    def tile_unsigned(a, _2q):
        out = []
        for x, q2 in zip(a, _2q):
            x.squeeze_()
            out.append(torch.from_numpy(O.C.tile_unsigned(_np(x), _np(q2))))
        return out

    m.mont_mult, m.mont_enter, m.ntt, m.enter_ntt = mont_mult, mont_enter, ntt, enter_ntt
    m.intt, m.intt_exit, m.intt_exit_reduce, m.intt_exit_reduce_signed = _intt(0), _intt(1), _intt(2), _intt(3)
    m.mont_redc, m.reduce_2q, m.make_signed, m.make_unsigned = mont_redc, reduce_2q, make_signed, make_unsigned
    m.mont_add, m.mont_sub, m.tile_unsigned = mont_add, mont_sub, tile_unsigned
    return m


sys.path.insert(0, str(REPO / "tests"))
from seeded_rng import SeededCsprng as CpuCsprng  # noqa: E402


_installed = {}


def install(cache_folder=None):
    """Make ``import liberate`` resolve to the reference sources with the stubs above."""
    if _installed:
        return _installed["cache"]
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, device=None, non_blocking=False, **k: self.to(device).clone()

    # package skeletons so that liberate/__init__.py (imports csprng .so's) is never executed
    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [str(path)]
        sys.modules[name] = m
        return m

    lib = pkg("liberate", REF_SRC / "liberate")
    cs = types.ModuleType("liberate.csprng")
    cs.Csprng = CpuCsprng
    sys.modules["liberate.csprng"] = cs
    lib.csprng = cs
    sys.modules["liberate.ntt.ntt_cuda"] = make_ntt_cuda_module()
    import importlib
    # same order as the reference's own liberate/__init__.py: fhe first (it pulls in liberate.ntt)
    lib.fhe = importlib.import_module("liberate.fhe")
    lib.ntt = importlib.import_module("liberate.ntt")

    # encdec.decode builds a device string "cpu:None" for CPU tensors (encdec.py:300); pre-seed its
    # caches under that key so the unmodified function runs on CPU.
    eng_mod = importlib.import_module("liberate.fhe.ckks_engine")
    encdec = importlib.import_module("liberate.fhe.encdec.encdec")
    _decode = encdec.decode

    def decode_cpu(m, *a, **k):
        N = len(m)
        key = (N, "cpu:None")
        if key not in encdec.perm_cache:
            encdec.perm_cache[key] = encdec.prepost_perms(N, device="cpu")
            encdec.skewer_cache[N, "cpu:None"] = encdec.generate_skewer(N, "cpu")
        return _decode(m, *a, **k)

    eng_mod.decode = decode_cpu

    if cache_folder is None:
        cache_folder = Path(tempfile.mkdtemp(prefix="refcache_"))
        for f in (REF_SRC / "liberate/fhe/cache/resources").glob("*.pkl"):
            shutil.copy(f, cache_folder / f.name)
    _installed["cache"] = str(cache_folder)
    return _installed["cache"]
