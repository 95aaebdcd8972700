"""
ntt_cuda -- drop-in for the reference's pybind extension ``liberate.ntt.ntt_cuda``
(src/liberate/ntt/ntt.cpp:421-437): the same 15 functions, the same argument lists
(``List[Tensor]``, one entry per participating device), the same in-place / returns-new behaviour.
Everything is forwarded to the C ABI of libckks_b200.so (include/ckks_b200.h) as raw device
pointers + row strides on the tensor's device and torch's current stream for that device.

Differences from the reference, none on the success path:
  * tensors must live on a CUDA device (the reference would crash in packed_accessor); no CPU path;
  * launch failures raise CkksLibError instead of being ignored (ntt.cpp checks nothing);
  * ``even``/``odd`` index tables are accepted and ignored; the painted ``psi[C, logN, N/2]`` table
    (ckks_context.py:336-341) is compacted once per tensor into ``[C, N]`` and cached.
"""
import weakref

import torch

from .._lib import lib, check


def _ptr(t):
    return t.data_ptr()


def _rows(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"ntt_cuda.{what}: tensor is on {t.device}; liberate_b200 has no CPU path")
    if t.dtype != torch.int64:
        raise TypeError(f"ntt_cuda.{what}: expected int64, got {t.dtype}")
    if t.dim() != 2 or (t.size(1) > 1 and t.stride(1) != 1):
        raise ValueError(f"ntt_cuda.{what}: expected a [C, N] tensor with contiguous rows, got "
                         f"shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0) if t.size(0) > 1 else t.size(1)


def _vec(t):
    if t.dim() != 1 or (t.numel() > 1 and t.stride(0) != 1):
        t = t.contiguous().view(-1)
    return t


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _Launch:
    """sets the tensor's device as current for the duration of the launch (kern.cu:103-107)"""

    def __init__(self, t):
        self.dev = t.device
        self.guard = None

    def __enter__(self):
        if torch.cuda.current_device() != self.dev.index:
            self.guard = torch.cuda.device(self.dev)
            self.guard.__enter__()
        return self

    def __exit__(self, *a):
        if self.guard is not None:
            self.guard.__exit__(*a)


_compact_cache = {}


def compact_twiddles(psi, forward):
    """painted psi[C, logN, N/2] -> compact [C, N]; cached per storage (keys die with the tensor)."""
    key = (psi.data_ptr(), tuple(psi.shape), tuple(psi.stride()), psi.device.index, bool(forward))
    hit = _compact_cache.get(key)
    if hit is not None and hit[0]() is not None:
        return hit[1]
    if psi.dim() == 2:  # already compact [C, N]
        return psi
    p = psi if psi.is_contiguous() else psi.contiguous()
    C, logN, half = p.shape
    out = torch.empty((C, 2 * half), dtype=torch.int64, device=p.device)
    with _Launch(p):
        check(lib.ckks_compact_twiddles(_ptr(p), _ptr(out), C, logN, 1 if forward else 0, _stream(p)), "compact_twiddles")
    base = psi._base if psi._base is not None else psi
    _compact_cache[key] = (weakref.ref(base), out)
    return out


# ---- the 15 operators ------------------------------------------------------------------------------
def mont_mult(a, b, ql, qh, kl, kh):
    outputs = []
    for x, y, l, h, k, kk in zip(a, b, ql, qh, kl, kh):
        xs, ys = _rows(x, "mont_mult"), _rows(y, "mont_mult")
        c = torch.empty_like(x, memory_format=torch.contiguous_format)
        with _Launch(x):
            check(lib.ckks_mont_mult(_ptr(x), xs, _ptr(y), ys, _ptr(c), c.size(1), x.size(0), x.size(1),
                                     _ptr(_vec(l)), _ptr(_vec(h)), _ptr(_vec(k)), _ptr(_vec(kk)), _stream(x)),
                  "mont_mult")
        outputs.append(c)
    return outputs


def mont_enter(a, Rs, ql, qh, kl, kh):
    for x, r, l, h, k, kk in zip(a, Rs, ql, qh, kl, kh):
        xs = _rows(x, "mont_enter")
        with _Launch(x):
            check(lib.ckks_mont_enter(_ptr(x), xs, _ptr(_vec(r)), x.size(0), x.size(1),
                                      _ptr(_vec(l)), _ptr(_vec(h)), _ptr(_vec(k)), _ptr(_vec(kk)), _stream(x)),
                  "mont_enter")


def _fwd(a, Rs, psi, _2q, ql, qh, kl, kh, what):
    for i, x in enumerate(a):
        xs = _rows(x, what)
        tw = compact_twiddles(psi[i], True)
        C = ql[i].size(0)
        logN = int(x.size(1)).bit_length() - 1
        rs = _ptr(_vec(Rs[i])) if Rs is not None else None
        with _Launch(x):
            check(lib.ckks_ntt(_ptr(x), xs, C, logN, _ptr(tw), tw.stride(0), rs, _ptr(_vec(_2q[i])),
                               _ptr(_vec(ql[i])), _ptr(_vec(qh[i])), _ptr(_vec(kl[i])), _ptr(_vec(kh[i])),
                               _stream(x)), what)


def ntt(a, even, odd, psi, _2q, ql, qh, kl, kh):
    _fwd(a, None, psi, _2q, ql, qh, kl, kh, "ntt")


def enter_ntt(a, Rs, even, odd, psi, _2q, ql, qh, kl, kh):
    _fwd(a, Rs, psi, _2q, ql, qh, kl, kh, "enter_ntt")


def _inv(a, psi, Ninv, _2q, ql, qh, kl, kh, mode, what):
    for i, x in enumerate(a):
        xs = _rows(x, what)
        tw = compact_twiddles(psi[i], False)
        C = ql[i].size(0)
        logN = int(x.size(1)).bit_length() - 1
        with _Launch(x):
            check(lib.ckks_intt(_ptr(x), xs, C, logN, _ptr(tw), tw.stride(0), _ptr(_vec(Ninv[i])),
                                _ptr(_vec(_2q[i])), _ptr(_vec(ql[i])), _ptr(_vec(qh[i])), _ptr(_vec(kl[i])),
                                _ptr(_vec(kh[i])), mode, _stream(x)), what)


def intt(a, even, odd, psi, Ninv, _2q, ql, qh, kl, kh):
    _inv(a, psi, Ninv, _2q, ql, qh, kl, kh, 0, "intt")


def intt_exit(a, even, odd, psi, Ninv, _2q, ql, qh, kl, kh):
    _inv(a, psi, Ninv, _2q, ql, qh, kl, kh, 1, "intt_exit")


def intt_exit_reduce(a, even, odd, psi, Ninv, _2q, ql, qh, kl, kh):
    _inv(a, psi, Ninv, _2q, ql, qh, kl, kh, 2, "intt_exit_reduce")


def intt_exit_reduce_signed(a, even, odd, psi, Ninv, _2q, ql, qh, kl, kh):
    _inv(a, psi, Ninv, _2q, ql, qh, kl, kh, 3, "intt_exit_reduce_signed")


def mont_redc(a, ql, qh, kl, kh):
    for x, l, h, k, kk in zip(a, ql, qh, kl, kh):
        xs = _rows(x, "mont_redc")
        with _Launch(x):
            check(lib.ckks_mont_redc(_ptr(x), xs, x.size(0), x.size(1), _ptr(_vec(l)), _ptr(_vec(h)),
                                     _ptr(_vec(k)), _ptr(_vec(kk)), _stream(x)), "mont_redc")


def _unary(fn, what):
    def op(a, _2q):
        for x, q2 in zip(a, _2q):
            xs = _rows(x, what)
            with _Launch(x):
                check(fn(_ptr(x), xs, x.size(0), x.size(1), _ptr(_vec(q2)), _stream(x)), what)
    op.__name__ = what
    return op


reduce_2q = _unary(lib.ckks_reduce_2q, "reduce_2q")
make_signed = _unary(lib.ckks_make_signed, "make_signed")
make_unsigned = _unary(lib.ckks_make_unsigned, "make_unsigned")


def _binary(fn, what):
    def op(a, b, _2q):
        outputs = []
        for x, y, q2 in zip(a, b, _2q):
            xs, ys = _rows(x, what), _rows(y, what)
            c = torch.empty_like(x, memory_format=torch.contiguous_format)
            with _Launch(x):
                check(fn(_ptr(x), xs, _ptr(y), ys, _ptr(c), c.size(1), x.size(0), x.size(1), _ptr(_vec(q2)),
                         _stream(x)), what)
            outputs.append(c)
        return outputs
    op.__name__ = what
    return op


mont_add = _binary(lib.ckks_mont_add, "mont_add")
mont_sub = _binary(lib.ckks_mont_sub, "mont_sub")


def tile_unsigned(a, _2q):
    outputs = []
    for x, q2 in zip(a, _2q):
        x.squeeze_()  # the reference squeezes its input in place (kern.cu:1206)
        if not x.is_cuda:
            raise RuntimeError("ntt_cuda.tile_unsigned: tensor is not on a CUDA device")
        src = x if x.is_contiguous() else x.contiguous()
        C, N = q2.size(0), src.size(0)
        c = torch.empty((C, N), dtype=x.dtype, device=x.device)
        with _Launch(x):
            check(lib.ckks_tile_unsigned(_ptr(src), _ptr(c), N, C, N, _ptr(_vec(q2)), _stream(x)), "tile_unsigned")
        outputs.append(c)
    return outputs
