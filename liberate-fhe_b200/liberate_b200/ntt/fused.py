"""
fused -- single-device wrappers of the level-2 entry points of libckks_b200.so
(include/ckks_b200.h): the reference's multi-launch Python sequences for rescale, tensor product,
Garner ModUp, evaluation-key inner product, ModDown and the Galois automorphism, each as one (or
two) kernels with the reference's exact per-element integer semantics.

All tensors: int64, CUDA, rows contiguous.  ``mp`` is a "mont pack" (_2q, ql, qh, kl, kh) of 1-D
per-limb tensors for the rows being processed.
"""
import torch

from .._lib import lib, check
from .ntt_cuda import _rows, _ptr, _stream, _Launch, _vec


def rescale(x, r0, scale, round_at, mp, canon=False):
    """engine.py:1026-1038.  x: [C,N] surviving limbs, r0: [N] dropped limb -> new [C,N]"""
    xs = _rows(x, "rescale")
    out = torch.empty((x.size(0), x.size(1)), dtype=torch.int64, device=x.device)
    r0 = r0 if r0.is_contiguous() else r0.contiguous()
    with _Launch(x):
        check(lib.ckks_rescale(_ptr(x), xs, _ptr(r0), _ptr(out), out.size(1), x.size(0), x.size(1), _ptr(_vec(scale)),
                               int(round_at), 1 if canon else 0, *[_ptr(_vec(t)) for t in mp], _stream(x)), "rescale")
    return out


def rescale_scaled(x, r0, scale, round_at, mp, pre=None, post=None):
    """rescale with mont_enter_scalar + reduce_2q folded in before (``pre``: mult_scalar, engine.py:2052-2098) and / or after
    it (``post``: level_up, engine.py:1410-1467): the integers of the kernel sequences, one pass.  r0: the dropped limb after
    the same pre-scaling.  -> new [C,N]"""
    xs = _rows(x, "rescale_scaled")
    out = torch.empty((x.size(0), x.size(1)), dtype=torch.int64, device=x.device)
    r0 = r0 if r0.is_contiguous() else r0.contiguous()
    opt = lambda t: _ptr(_vec(t)) if t is not None else None
    with _Launch(x):
        check(lib.ckks_rescale_scaled(_ptr(x), xs, _ptr(r0), _ptr(out), out.size(1), x.size(0), x.size(1), opt(pre),
                                      _ptr(_vec(scale)), int(round_at), opt(post), *[_ptr(_vec(t)) for t in mp], _stream(x)),
              "rescale_scaled")
    return out


def pc_product(p, c0, c1, mp, out0, out1):
    """plaintext x ciphertext in the NTT domain (mc_mult, engine.py:2100-2140): out0 = mont(p, c0), out1 = mont(p, c1);
    p, c0, c1 share one row stride; out may alias c"""
    s = _rows(p, "pc_product")
    if _rows(c0, "pc_product") != s or _rows(c1, "pc_product") != s or _rows(out0, "pc_product") != _rows(out1, "pc_product"):
        raise ValueError("pc_product: operands must share one row stride")
    C, N = p.shape
    with _Launch(p):
        check(lib.ckks_pc_product(_ptr(p), _ptr(c0), _ptr(c1), s, _ptr(out0), _ptr(out1), _rows(out0, "pc_product"), C, N,
                                  *[_ptr(_vec(t)) for t in mp], _stream(p)), "pc_product")


def pc_add(p, c0, Rs_scale, Rs, mp):
    """plaintext + ciphertext (mc_add, engine.py:2142-2175: mont_enter_scale, mont_enter, mont_add, mont_redc, reduce_2q)
    in one pass -> new [C,N]"""
    sp, sc = _rows(p, "pc_add"), _rows(c0, "pc_add")
    out = torch.empty((c0.size(0), c0.size(1)), dtype=torch.int64, device=c0.device)
    with _Launch(c0):
        check(lib.ckks_pc_add(_ptr(p), sp, _ptr(c0), sc, _ptr(out), out.size(1), c0.size(0), c0.size(1), _ptr(_vec(Rs_scale)),
                              _ptr(_vec(Rs)), *[_ptr(_vec(t)) for t in mp], _stream(c0)), "pc_add")
    return out


def addsub_reduce(a, b, _2q, sub=False):
    """mont_add / mont_sub + reduce_2q in one pass (cc_add / cc_sub, engine.py:1268-1330) -> new [C,N]"""
    sa, sb = _rows(a, "addsub_reduce"), _rows(b, "addsub_reduce")
    out = torch.empty((a.size(0), a.size(1)), dtype=torch.int64, device=a.device)
    with _Launch(a):
        check(lib.ckks_addsub_reduce(_ptr(a), sa, _ptr(b), sb, _ptr(out), out.size(1), a.size(0), a.size(1), _ptr(_vec(_2q)),
                                     1 if sub else 0, _stream(a)), "addsub_reduce")
    return out


def tensor_product(x0, x1, y0, y1, mp):
    """engine.py:1095-1101 -> (d0, d1, d2)"""
    s = _rows(x0, "tensor_product")
    for t in (x1, y0, y1):
        if _rows(t, "tensor_product") != s:
            raise ValueError("tensor_product: operands must share one row stride")
    C, N = x0.shape
    d = torch.empty((3, C, N), dtype=torch.int64, device=x0.device)
    with _Launch(x0):
        check(lib.ckks_tensor_product(_ptr(x0), _ptr(x1), _ptr(y0), _ptr(y1), s, _ptr(d[0]), _ptr(d[1]), _ptr(d[2]),
                                      N, C, N, *[_ptr(_vec(t)) for t in mp], _stream(x0)), "tensor_product")
    return d[0], d[1], d[2]


def garner_digits(a_part, Y_scalar, Ltri, mp4, out=None):
    """pre_extend, engine.py:654-705.  a_part: [alpha,N] -> state [alpha,N] (plain integers)"""
    s = _rows(a_part, "garner_digits")
    alpha, N = a_part.shape
    if out is None:
        out = torch.empty((alpha, N), dtype=torch.int64, device=a_part.device)
    with _Launch(a_part):
        check(lib.ckks_garner_digits(_ptr(a_part), s, _ptr(out), _rows(out, "garner_digits"), alpha, N,
                                     _ptr(Y_scalar) if Y_scalar is not None else None,
                                     _ptr(Ltri) if Ltri is not None else None,
                                     *[_ptr(_vec(t)) for t in mp4], _stream(a_part)), "garner_digits")
    return out


def extend(state, Rs, Lenter, mp, out=None, canon=False):
    """extend, engine.py:707-743.  state [alpha,N] -> [E,N] Montgomery form on the E target limbs"""
    s = _rows(state, "extend")
    alpha, N = state.shape
    E = Rs.numel()
    if out is None:
        out = torch.empty((E, N), dtype=torch.int64, device=state.device)
    with _Launch(state):
        check(lib.ckks_extend(_ptr(state), s, alpha, _ptr(out), _rows(out, "extend"), E, N, _ptr(_vec(Rs)),
                              _ptr(Lenter) if Lenter is not None else None, 1 if canon else 0,
                              *[_ptr(_vec(t)) for t in mp], _stream(state)), "extend")
    return out


def ksk_accumulate(ext, ksk0, ksk1, acc0, acc1, first, mp):
    """engine.py:906-937 + 832-840: acc_i (+)= ext (*) ksk_i"""
    E, N = ext.shape
    ks = _rows(ksk0, "ksk_accumulate")
    if _rows(ksk1, "ksk_accumulate") != ks:
        raise ValueError("ksk_accumulate: key halves must share one row stride")
    with _Launch(ext):
        check(lib.ckks_ksk_accumulate(_ptr(ext), _rows(ext, "ksk"), _ptr(ksk0), _ptr(ksk1), ks, _ptr(acc0), _ptr(acc1),
                                      _rows(acc0, "ksk"), E, N, 1 if first else 0,
                                      *[_ptr(_vec(t)) for t in mp], _stream(ext)), "ksk_accumulate")


def moddown(d, L, K, Rs, PiR, mp, add=None, eff=None):
    """engine.py:851-901 (+ the add/reduce tail of relinearize / switch_key) -> new [L,N]"""
    E, N = d.shape
    assert E == L + K
    out = torch.empty((L, N), dtype=torch.int64, device=d.device)
    if eff is None:
        eff = torch.empty((K, N), dtype=torch.int64, device=d.device)
    with _Launch(d):
        check(lib.ckks_moddown(_ptr(d), _rows(d, "moddown"), L, K, N, _ptr(_vec(Rs)), _ptr(PiR),
                               _ptr(add) if add is not None else None,
                               _rows(add, "moddown") if add is not None else 0,
                               _ptr(out), N, _ptr(eff), *[_ptr(_vec(t)) for t in mp], _stream(d)), "moddown")
    return out


def automorphism(x, g, canon, _2q=None):
    """encdec.rotate/conjugate (encdec.py:224-270) [+ make_unsigned + reduce_2q, engine.py:1196-1200]"""
    s = _rows(x, "automorphism")
    C, N = x.shape
    out = torch.empty((C, N), dtype=torch.int64, device=x.device)
    with _Launch(x):
        check(lib.ckks_automorphism(_ptr(x), s, _ptr(out), N, C, N, int(g), 1 if canon else 0,
                                    _ptr(_vec(_2q)) if _2q is not None else None, _stream(x)), "automorphism")
    return out


class FastTables:
    """fast-transform tables of one direction for the limbs of one device: plain {w, w'} / double tables plus the
    PACKED copies of the last four stages (include/ckks_b200.h: ckks_fast_pack).  Slicing keeps the four in step."""

    def __init__(self, sh, dbl, psh, pdbl):
        self.sh, self.dbl, self.psh, self.pdbl = sh, dbl, psh, pdbl

    def __getitem__(self, sl):
        return FastTables(self.sh[sl], self.dbl[sl], self.psh[sl], self.pdbl[sl])

    def __iter__(self):             # (sh, dbl) = tables: the two plain tables, as before
        return iter((self.sh, self.dbl))


def fast_tables(plain, q):
    """plain canonical twiddles [C,N] -> FastTables (shoup [C,N,2] int64 {w, floor(w 2^64/q)}, dbl [C,N] float64, packed copies)"""
    C, N = plain.shape
    logN = int(N).bit_length() - 1
    sh = torch.empty((C, N, 2), dtype=torch.int64, device=plain.device)
    dbl = torch.empty((C, N), dtype=torch.float64, device=plain.device)
    psh, pdbl = torch.empty_like(sh), torch.empty_like(dbl)
    with _Launch(plain):
        check(lib.ckks_fast_tables(_ptr(plain), _ptr(_vec(q)), _ptr(sh), _ptr(dbl), C, N, _stream(plain)), "fast_tables")
        check(lib.ckks_fast_pack(_ptr(sh), _ptr(dbl), _ptr(psh), _ptr(pdbl), C, logN, _stream(plain)), "fast_pack")
    return FastTables(sh, dbl, psh, pdbl)


def reciprocals(q):
    """[C] float64 1/q, correctly rounded (what the kernels would compute per CTA with an FP64 division)"""
    return torch.tensor([1.0 / float(int(x)) for x in q.tolist()], dtype=torch.float64, device=q.device)


def _tables(tables, tw_f64):
    """accept FastTables or the (tw_u64, tw_f64) pair -> the four table pointers"""
    if isinstance(tables, FastTables):
        return _ptr(tables.sh), _ptr(tables.dbl), _ptr(tables.psh), _ptr(tables.pdbl)
    return _ptr(tables), _ptr(tw_f64), None, None


def ntt_fast(x, tables, tw_f64, q, scal=None, scal_sh=None, period=None, force_int=False, qinv=None, perm=False):
    """canonical-output forward NTT in place: x [rows,N] in [0,2q) -> NTT(x * scal) in [0,q).
    tables: FastTables (tw_f64 ignored) or the plain {w, w'} table with tw_f64 the double table"""
    s = _rows(x, "ntt_fast")
    rows, N = x.shape
    logN = int(N).bit_length() - 1
    with _Launch(x):
        check(lib.ckks_ntt_fast(_ptr(x), s, rows, period or rows, logN, *_tables(tables, tw_f64), _ptr(_vec(q)),
                                _ptr(qinv) if qinv is not None else None,
                                _ptr(scal) if scal is not None else None,
                                _ptr(scal_sh) if scal_sh is not None else None, 1 if force_int else 0, 1 if perm else 0,
                                _stream(x)), "ntt_fast")


def intt_fast(x, tables, tw_f64, q, scal, scal_sh, centred=False, period=None, force_int=False, qinv=None, perm=False):
    """canonical-output inverse NTT in place: x [rows,N] in [0,2q) -> iNTT(x) * scal in [0,q) (or centred)"""
    s = _rows(x, "intt_fast")
    rows, N = x.shape
    logN = int(N).bit_length() - 1
    with _Launch(x):
        check(lib.ckks_intt_fast(_ptr(x), s, rows, period or rows, logN, *_tables(tables, tw_f64), _ptr(_vec(q)),
                                 _ptr(qinv) if qinv is not None else None,
                                 _ptr(scal), _ptr(scal_sh), 1 if centred else 0, 1 if force_int else 0, 1 if perm else 0,
                                 _stream(x)), "intt_fast")


def perm_rows(x, inverse=False):
    """NTT-domain rows natural -> warp-interleaved order of the fused executor (inverse=True: back); returns a new tensor"""
    s = _rows(x, "perm_rows")
    rows, N = x.shape
    out = torch.empty((rows, N), dtype=torch.int64, device=x.device)
    with _Launch(x):
        check(lib.ckks_perm_rows(_ptr(x), s, _ptr(out), N, rows, N, 1 if inverse else 0, _stream(x)), "perm_rows")
    return out


def ksk_inner(ext_all, parts, k0_ptrs, k1_ptrs, ksk_stride, acc0, acc1, mp):
    """one-pass evaluation-key inner product over all partitions; ext_all: [parts*E, N]; k*_ptrs: int64 device
    tensors holding `parts` raw row-0 pointers of the key halves"""
    E, N = acc0.shape
    with _Launch(ext_all):
        check(lib.ckks_ksk_inner(_ptr(ext_all), _rows(ext_all, "ksk_inner"), parts, _ptr(k0_ptrs), _ptr(k1_ptrs),
                                 ksk_stride, _ptr(acc0), _ptr(acc1), _rows(acc0, "ksk_inner"), E, N,
                                 *[_ptr(_vec(t)) for t in mp], _stream(ext_all)), "ksk_inner")
