"""
ntt_context -- per-device parameter tables and the thin operator wrappers the engine calls.

API of the reference class (src/liberate/ntt/ntt_context.py:13-599) that the engine and user code touch is
kept: ``devices, num_devices, p, starts, stops, q, _2q, ql, qh, kl, kh, Rs, Rs_scale, Ninv, psi, ipsi, qlists,
parts_pack[dev][key]{Y_scalar, L_scalar, L_enter}`` and the wrappers ``mont_enter / mont_enter_scale /
mont_enter_scalar / mont_mult / ntt / enter_ntt / intt / mont_redc / intt_exit / intt_exit_reduce /
intt_exit_reduce_signed / reduce_2q / make_signed / make_unsigned / mont_add / mont_sub / tile_unsigned``
with the reference's ``(a, lvl=0, mult_type=-1, part=0)`` addressing.

What changed underneath (B200-first):
  * twiddles are a compact ``[limbs, N]`` table per direction, grown ON THE GPU by logN doubling steps from
    per-prime scalars (ckks_context.psi_stage_factors) and then taken to Montgomery form by the same
    mont_enter the reference uses (nctx.py:115-130) -- so the values are bit-identical to the reference's
    painted psi[C, logN, N/2] at 1/logN of the memory, with no Python-int power series and no pickle cache;
  * there are no even/odd index tables;
  * the (level, mult_type, part) -> row-range addressing is computed on demand instead of being
    pre-materialised as nested lists of views (nctx.py:417-526);
  * one process may own only some of the logical devices (``local_ids``): the one-process-per-GPU
    torch.distributed layout.  Lists stay indexed by logical device id; non-local entries are None.
"""
import numpy as np
import torch

from ..fhe.presets import errors
from . import fused, ntt_cuda
from .rns_partition import rns_partition


@errors.log_error
class ntt_context:
    def __init__(self, ctx, index_type=torch.int32, devices=None, verbose=False, local_ids=None):
        if devices is None:
            devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
            if not devices:
                raise errors.DeviceSelectError()
        self.devices = [f"cuda:{d}" if isinstance(d, int) else d for d in devices]
        self.num_devices = len(self.devices)
        self.local_ids = list(range(self.num_devices)) if local_ids is None else list(local_ids)
        self.index_type = index_type
        self.verbose = verbose
        self.ctx = ctx
        self.num_ordinary_primes = ctx.num_scales + 1
        self.num_special_primes = ctx.num_special_primes
        self.num_levels = ctx.num_scales + 1
        self.p = rns_partition(self.num_ordinary_primes, self.num_special_primes, self.num_devices)

        self.starts = self.p.diff
        self.stops = [[len(d) for d in self.p.destination_arrays_with_special[0]],
                      [len(d) for d in self.p.destination_arrays[0]]]
        self._tables()
        self.qlists = [[ctx.q[i] for i in rows] for rows in self.p.d_special]   # host ints, every logical device
        self._garner_tables()
        self._cache = {}

    def is_local(self, dev_id):
        return dev_id in self.local_ids

    # ---------------------------------------------------------------------------------------------
    # device tables
    # ---------------------------------------------------------------------------------------------
    def partition_variable(self, variable):
        """per-device slices (rows in the device's prime order, specials last), nctx.py:95-107"""
        v = np.array(variable, dtype=np.int64)
        return [torch.from_numpy(v[self.p.d_special[d]]).to(self.devices[d]) if self.is_local(d) else None
                for d in range(self.num_devices)]

    def _tables(self):
        c = self.ctx
        R = c.R
        scale = 2 ** c.scale_bits
        self.Rs = self.partition_variable(c.R_square)
        self.Rs_scale = self.partition_variable([(rs * scale) % q for rs, q in zip(c.R_square, c.q)])
        self.q = self.partition_variable(c.q)
        self._2q = self.partition_variable(c.q_double)
        self.ql = self.partition_variable(c.q_lower_bits)
        self.qh = self.partition_variable(c.q_higher_bits)
        self.kl = self.partition_variable(c.k_lower_bits)
        self.kh = self.partition_variable(c.k_higher_bits)
        self.Ninv = self.partition_variable([(ni * R) % q for ni, q in zip(c.N_inv, c.q)])
        self.mont_pack0 = [self.ql, self.qh, self.kl, self.kh]
        fwd, inv = c.psi_stage_factors()
        # plain-twiddle tables of the canonical-output fast transforms (fused.FastTables: {w, floor(w 2^64/q)} [rows,N,2],
        # double [rows,N], and the packed copies of the last four stages)
        self.tw_fast_fwd = [None] * self.num_devices
        self.tw_fast_inv = [None] * self.num_devices
        self.psi = [self._grow_twiddles(d, fwd, self.tw_fast_fwd) if self.is_local(d) else None
                    for d in range(self.num_devices)]
        self.ipsi = [self._grow_twiddles(d, inv, self.tw_fast_inv) if self.is_local(d) else None
                     for d in range(self.num_devices)]
        # per-limb plain scalars of the fast transforms with their Shoup companions (uint64 bit patterns)
        Rinv = [pow(R, -1, q) for q in c.q]
        self.qinv = [fused.reciprocals(self.q[d]) if self.is_local(d) else None for d in range(self.num_devices)]
        self.fs_R = self._fast_scalar([R % q for q in c.q])                                   # "enter": x R
        self.fs_exit = self._fast_scalar([ni * ri % q for ni, ri, q in zip(c.N_inv, Rinv, c.q)])  # x N^-1 R^-1
        self.fs_ninv = self._fast_scalar(list(c.N_inv))                                        # x N^-1

    def _fast_scalar(self, values):
        def pattern(v, q):
            w = (v << 64) // q
            return w - (1 << 64) if w >= (1 << 63) else w
        sh = [pattern(v, q) for v, q in zip(values, self.ctx.q)]
        return self.partition_variable(values), self.partition_variable(sh)

    def _grow_twiddles(self, dev, factors, fast_store):
        """compact bit-reversed power table [limbs, N] in Montgomery form for device ``dev``.
        table[m + i] = table[i] * w^(N/2m), m = 2^s (see ckks_context.psi_stage_factors)."""
        c = self.ctx
        rows = self.p.d_special[dev]
        N, logN, R = c.N, c.logN, c.R
        device = self.devices[dev]
        SEED = 16  # first entries on the host so that every device-side slice is 16-byte aligned
        seed = np.zeros((len(rows), SEED), dtype=np.int64)
        for r, pi in enumerate(rows):
            q = c.q[pi]
            vals = [1]
            for s in range(4):
                f = factors[pi][s]
                vals = vals + [v * f % q for v in vals]
            seed[r] = [v * R % q for v in vals]       # Montgomery form, canonical
        T = torch.zeros((len(rows), N), dtype=torch.int64, device=device)
        T[:, :SEED] = torch.from_numpy(seed).to(device)
        mont = [[t[dev]] for t in self.mont_pack0]
        for s in range(4, logN):
            m = 1 << s
            step = torch.tensor([factors[pi][s] * R % c.q[pi] for pi in rows], dtype=torch.int64, device=device)
            T[:, m:2 * m] = T[:, :m]
            ntt_cuda.mont_enter([T[:, m:2 * m]], [step], *mont)
        # canonical plain values, then the reference's own Montgomery entry (nctx.py:115-130)
        ntt_cuda.mont_redc([T], *mont)
        ntt_cuda.reduce_2q([T], [self._2q[dev]])
        fast_store[dev] = fused.fast_tables(T, self.q[dev])
        ntt_cuda.mont_enter([T], [self.Rs[dev]], *mont)
        return T

    # ---------------------------------------------------------------------------------------------
    # Garner (mixed-radix ModUp) constants, ntt_context.py:323-412
    # ---------------------------------------------------------------------------------------------
    def _garner_tables(self):
        c = self.ctx
        R = c.R
        self.parts_pack = [dict() for _ in range(self.num_devices)]
        for dev in range(self.num_devices):
            for level in range(self.num_levels):
                for part_index, part in enumerate(self.p.destination_parts[level][dev]):
                    key = tuple(self.p.p[level][dev][part_index])
                    if key in self.parts_pack[dev]:
                        continue
                    m = [c.q[i] for i in part]
                    alpha = len(m)
                    Lprod = [m[0]]
                    for i in range(1, alpha - 1):
                        Lprod.append(Lprod[-1] * m[i])
                    Y = [pow(Lprod[i], -1, m[i + 1]) * R % m[i + 1] for i in range(alpha - 1)]
                    Ls = [[Lprod[i] * R % m[j] for j in range(i + 2, alpha)] for i in range(alpha - 2)]
                    L_enter = []
                    for tgt in range(self.num_devices):
                        dest = self.p.destination_arrays_with_special[0][tgt]
                        L_enter.append([[Lprod[i] * c.R_square[j] % c.q[j] for j in dest] for i in range(alpha - 1)])
                    self.parts_pack[dev][key] = dict(prime_ids=list(part), alpha=alpha, Y_host=Y, L_host=Ls,
                                                     L_enter_host=L_enter)

    def garner(self, level, dev, part_index):
        """device tensors for the fused digit kernel of (level, source device, part): Y [alpha-1],
        Ltri [(alpha-1), alpha] (entry (i, j) = L_scalar[i][j-(i+2)]), mont pack of the alpha rows"""
        key = tuple(self.p.p[level][dev][part_index])
        item = self.parts_pack[dev][key]
        if "Y_scalar" not in item:
            device = self.devices[dev]
            alpha = item["alpha"]
            t = lambda v: torch.tensor(v, dtype=torch.int64, device=device)
            item["Y_scalar"] = t(item["Y_host"]) if alpha > 1 else None
            item["L_scalar"] = [t(row) for row in item["L_host"]] if alpha > 2 else None
            tri = np.zeros((max(alpha - 1, 1), alpha), dtype=np.int64)
            for i, row in enumerate(item["L_host"]):
                tri[i, i + 2:i + 2 + len(row)] = row
            item["Ltri"] = torch.from_numpy(tri).to(device) if alpha > 2 else None
            a, b = key[0], key[-1] + 1
            item["mont4"] = [p[dev][a:b] for p in self.mont_pack0]
        return item

    def lenter(self, level, src_dev, part_index, dst_dev):
        """contiguous [(alpha-1), E] table of (L_i * R^2) mod q_t for the rows of dst_dev live at `level`"""
        ck = ("lenter", level, src_dev, part_index, dst_dev)
        hit = self._cache.get(ck)
        if hit is None:
            key = tuple(self.p.p[level][src_dev][part_index])
            item = self.parts_pack[src_dev][key]
            if item["alpha"] == 1:
                hit = (None,)
            else:
                start = self.starts[level][dst_dev]
                rows = [r[start:] for r in item["L_enter_host"][dst_dev]]
                hit = (torch.tensor(rows, dtype=torch.int64, device=self.devices[dst_dev]),)
            self._cache[ck] = hit
        return hit[0]

    # ---------------------------------------------------------------------------------------------
    # (level, mult_type, part) -> row ranges, nctx.py:417-526 in closed form
    # ---------------------------------------------------------------------------------------------
    def rows(self, lvl, mult_type, part):
        """list of (device id, first row, one-past-last row) in level-0 row numbering"""
        if mult_type < 0:
            stops = self.stops[mult_type]           # -2 -> with special (index 0), -1 -> ordinary (index 1)
            out = [(d, self.starts[lvl][d], stops[d]) for d in range(self.num_devices)]
            return [(d, a, b) for d, a, b in out if b > a]
        d = mult_type
        if part < 0:
            a, b = self.starts[lvl][d], self.stops[part][d]
        else:
            rows = self.p.p_special[lvl][d][part]
            a, b = rows[0], rows[-1] + 1
        return [(d, a, b)] if b > a else []

    def _sel(self, table, lvl, mult_type, part):
        return [table[d][a:b] for d, a, b in self.rows(lvl, mult_type, part) if self.is_local(d)]

    def mont_pack(self, lvl, mult_type, part):
        return [self._sel(t, lvl, mult_type, part) for t in self.mont_pack0]

    def pack5(self, lvl, dev, part=-2):
        """(_2q, ql, qh, kl, kh) row slices of one device, for the fused single-device operators"""
        (_, a, b), = self.rows(lvl, dev, part)
        return [t[dev][a:b] for t in (self._2q, self.ql, self.qh, self.kl, self.kh)]

    # ---------------------------------------------------------------------------------------------
    # canonical-output fast transforms on one device's tensor (fused path)
    # ---------------------------------------------------------------------------------------------
    def ntt_fast(self, x, lvl, dev, part=-1, enter=False, batched=False):
        """x: [rows, N] (or [parts*rows, N] with batched=True) in [0,2q) -> canonical NTT(x [* R])"""
        (_, a, b), = self.rows(lvl, dev, part)
        sc = (self.fs_R[0][dev][a:b], self.fs_R[1][dev][a:b]) if enter else (None, None)
        fused.ntt_fast(x, self.tw_fast_fwd[dev][a:b], None, self.q[dev][a:b], sc[0], sc[1],
                       period=(b - a) if batched else None, qinv=self.qinv[dev][a:b])

    def intt_fast(self, x, lvl, dev, part=-1, exit=True, centred=False, batched=False):
        """x: [rows, N] (or [parts*rows, N] with batched=True) in [0,2q) -> canonical iNTT(x) * N^-1 [* R^-1]
        (== intt_exit_reduce / intt + reduce)"""
        (_, a, b), = self.rows(lvl, dev, part)
        sc = self.fs_exit if exit else self.fs_ninv
        fused.intt_fast(x, self.tw_fast_inv[dev][a:b], None, self.q[dev][a:b], sc[0][dev][a:b], sc[1][dev][a:b],
                        centred=centred, period=(b - a) if batched else None, qinv=self.qinv[dev][a:b])

    @staticmethod
    def _live(a):
        return [x for x in a if x is not None]

    @staticmethod
    def _expand(like, outs):
        it = iter(outs)
        return [None if x is None else next(it, None) for x in like]

    # ---------------------------------------------------------------------------------------------
    # operator wrappers (nctx.py:532-599)
    # ---------------------------------------------------------------------------------------------
    def mont_enter(self, a, lvl=0, mult_type=-1, part=0):
        ntt_cuda.mont_enter(self._live(a), self._sel(self.Rs, lvl, mult_type, part), *self.mont_pack(lvl, mult_type, part))

    def mont_enter_scale(self, a, lvl=0, mult_type=-1, part=0):
        ntt_cuda.mont_enter(self._live(a), self._sel(self.Rs_scale, lvl, mult_type, part),
                            *self.mont_pack(lvl, mult_type, part))

    def mont_enter_scalar(self, a, b, lvl=0, mult_type=-1, part=0):
        ntt_cuda.mont_enter(self._live(a), self._live(b), *self.mont_pack(lvl, mult_type, part))

    def mont_mult(self, a, b, lvl=0, mult_type=-1, part=0):
        return self._expand(a, ntt_cuda.mont_mult(self._live(a), self._live(b), *self.mont_pack(lvl, mult_type, part)))

    def _tw(self, table, lvl, mult_type, part):
        return self._sel(table, lvl, mult_type, part)

    def ntt(self, a, lvl=0, mult_type=-1, part=0):
        ntt_cuda.ntt(self._live(a), None, None, self._tw(self.psi, lvl, mult_type, part),
                     self._sel(self._2q, lvl, mult_type, part), *self.mont_pack(lvl, mult_type, part))

    def enter_ntt(self, a, lvl=0, mult_type=-1, part=0):
        ntt_cuda.enter_ntt(self._live(a), self._sel(self.Rs, lvl, mult_type, part), None, None,
                           self._tw(self.psi, lvl, mult_type, part), self._sel(self._2q, lvl, mult_type, part),
                           *self.mont_pack(lvl, mult_type, part))

    def _inv(self, fn, a, lvl, mult_type, part):
        fn(self._live(a), None, None, self._tw(self.ipsi, lvl, mult_type, part),
           self._sel(self.Ninv, lvl, mult_type, part), self._sel(self._2q, lvl, mult_type, part),
           *self.mont_pack(lvl, mult_type, part))

    def intt(self, a, lvl=0, mult_type=-1, part=0):
        self._inv(ntt_cuda.intt, a, lvl, mult_type, part)

    def intt_exit(self, a, lvl=0, mult_type=-1, part=0):
        self._inv(ntt_cuda.intt_exit, a, lvl, mult_type, part)

    def intt_exit_reduce(self, a, lvl=0, mult_type=-1, part=0):
        self._inv(ntt_cuda.intt_exit_reduce, a, lvl, mult_type, part)

    def intt_exit_reduce_signed(self, a, lvl=0, mult_type=-1, part=0):
        self._inv(ntt_cuda.intt_exit_reduce_signed, a, lvl, mult_type, part)

    def mont_redc(self, a, lvl=0, mult_type=-1, part=0):
        ntt_cuda.mont_redc(self._live(a), *self.mont_pack(lvl, mult_type, part))

    def reduce_2q(self, a, lvl=0, mult_type=-1, part=0):
        ntt_cuda.reduce_2q(self._live(a), self._sel(self._2q, lvl, mult_type, part))

    def make_signed(self, a, lvl=0, mult_type=-1, part=0):
        ntt_cuda.make_signed(self._live(a), self._sel(self._2q, lvl, mult_type, part))

    def make_unsigned(self, a, lvl=0, mult_type=-1, part=0):
        ntt_cuda.make_unsigned(self._live(a), self._sel(self._2q, lvl, mult_type, part))

    def mont_add(self, a, b, lvl=0, mult_type=-1, part=0):
        return self._expand(a, ntt_cuda.mont_add(self._live(a), self._live(b), self._sel(self._2q, lvl, mult_type, part)))

    def mont_sub(self, a, b, lvl=0, mult_type=-1, part=0):
        return self._expand(a, ntt_cuda.mont_sub(self._live(a), self._live(b), self._sel(self._2q, lvl, mult_type, part)))

    def addsub_reduce(self, a, b, sub, lvl=0, mult_type=-1, part=0):
        """mont_add / mont_sub followed by reduce_2q as ONE kernel per device (same integers as the two-call sequence)"""
        q2 = self._sel(self._2q, lvl, mult_type, part)
        outs = [fused.addsub_reduce(x, y, q, sub) for x, y, q in zip(self._live(a), self._live(b), q2)]
        return self._expand(a, outs)

    def tile_unsigned(self, a, lvl=0, mult_type=-1, part=0):
        return self._expand(a, ntt_cuda.tile_unsigned(self._live(a), self._sel(self._2q, lvl, mult_type, part)))
