"""
executor -- host side of the level-3 C ABI (include/ckks_b200.h: ckks_exec_tensor_stage / ckks_exec_digits /
ckks_exec_keyswitch_stage).  One ctypes call launches a whole stage of the hot path (10-12 kernels), so the Python
cost of a ct x ct multiplication drops from ~50 operator calls to 2.

A ``LevelPlan`` is the (level, device) descriptor the C side reads: row constants, fast-transform tables, Garner /
ModDown tables, partition bookkeeping -- all device pointers into tensors owned by ntt_context / the engine, plus
PERSISTENT workspaces (so that every pointer the kernels see is constant from call to call: nothing is allocated
and no pointer table is uploaded on the hot path).
"""
import ctypes

import torch

from .._lib import check, lib

_vp = ctypes.c_void_p


class LevelT(ctypes.Structure):
    _fields_ = [("logN", ctypes.c_int32), ("L", ctypes.c_int32), ("K", ctypes.c_int32), ("nparts", ctypes.c_int32),
                ("nlocal", ctypes.c_int32), ("_pad", ctypes.c_int32),
                ("q", _vp), ("_2q", _vp), ("ql", _vp), ("qh", _vp), ("kl", _vp), ("kh", _vp), ("Rs", _vp),
                ("twf_u64", _vp), ("twf_f64", _vp), ("twi_u64", _vp), ("twi_f64", _vp),
                ("sR", _vp), ("sR_sh", _vp), ("sExit", _vp), ("sExit_sh", _vp),
                ("PiR", _vp), ("part_alpha", _vp), ("Lenter", _vp),
                ("loc_row0", _vp), ("loc_alpha", _vp), ("loc_Y", _vp), ("loc_Ltri", _vp),
                ("rescale_scale", _vp), ("round_at", ctypes.c_int64),
                ("part_wide", _vp), ("Hm", _vp), ("Rd", _vp), ("C31", _vp),
                ("Rinv", _vp), ("Pinv", _vp), ("L_small", ctypes.c_int32), ("amax", ctypes.c_int32),
                ("twpf_u64", _vp), ("twpf_f64", _vp), ("twpi_u64", _vp), ("twpi_f64", _vp), ("qinv", _vp)]


MAX_ALPHA = 8        # limbs per key-switch partition the fused kernels support (csrc/ckks_b200.cu)


def _p(t):
    return t.data_ptr() if t is not None else None


class Workspace:
    """per-device scratch, sized once for the largest level in use and then sliced (constant addresses)"""

    def __init__(self, device):
        self.device = device
        self.buf = {}

    def get(self, name, elems):
        t = self.buf.get(name)
        if t is None or t.numel() < elems:
            t = torch.empty(elems, dtype=torch.int64, device=self.device)
            self.buf[name] = t
        return t[:elems]


class LevelPlan:
    def __init__(self, eng, level, dev):
        ntt = eng.ntt
        self.level, self.dev = level, dev
        device = ntt.devices[dev]
        N = eng.ctx.N
        K = ntt.num_special_primes
        (_, a, b), = ntt.rows(level, dev, -2)
        E = b - a
        self.L, self.K, self.E, self.N = E - K, K, E, N
        keep = []                      # tensors the descriptor points into

        def hold(t):
            keep.append(t)
            return t

        owners = eng._part_owners(level)
        # partition order of this plan: the partitions whose digits this device makes first (their extension and transforms
        # can start before the peers' digits have arrived), then the others; sums over partitions do not depend on the order
        self.sids = sorted(owners, key=lambda s_: (owners[s_][0] != dev, s_))
        self.owners = owners
        if max(o[2] for o in owners.values()) > MAX_ALPHA:
            # (num_special_primes > 8 makes partitions of more than 8 limbs: the kernels would silently drop digits)
            raise NotImplementedError(f"key-switch partitions of more than {MAX_ALPHA} limbs are not supported by the "
                                      "fused path; construct the engine with fast=False")
        i32 = lambda v: hold(torch.tensor(v, dtype=torch.int32, device=device))
        i64 = lambda v: hold(torch.tensor(v, dtype=torch.int64, device=device))
        part_alpha = i32([owners[s][2] for s in self.sids])
        lenter = [ntt.lenter(level, owners[s][0], owners[s][1], dev) for s in self.sids]
        keep.extend(t for t in lenter if t is not None)
        lenter_ptrs = i64([_p(t) or 0 for t in lenter])
        local = [(s, owners[s]) for s in self.sids if owners[s][0] == dev]
        self.local_sids = [s for s, _ in local]
        g = [ntt.garner(level, dev, o[1]) for _, o in local]
        self.loc_row0 = [ntt.p.parts[level][dev][o[1]][0] for _, o in local]
        self.loc_alpha = [o[2] for _, o in local]
        loc_row0 = i32(self.loc_row0 or [0])
        loc_alpha = i32(self.loc_alpha or [0])
        loc_Y = i64([_p(x["Y_scalar"]) or 0 for x in g] or [0])
        loc_L = i64([_p(x["Ltri"]) or 0 for x in g] or [0])

        tf, ti = ntt.tw_fast_fwd[dev], ntt.tw_fast_inv[dev]
        sl = slice(a, b)
        T = lambda t: hold(t[dev][sl])
        d = LevelT()
        d.logN, d.L, d.K, d.nparts, d.nlocal = eng.ctx.logN, self.L, K, len(self.sids), len(local)
        d.q, d._2q, d.ql, d.qh, d.kl, d.kh, d.Rs = (_p(T(x)) for x in (ntt.q, ntt._2q, ntt.ql, ntt.qh, ntt.kl, ntt.kh, ntt.Rs))
        d.twf_u64, d.twf_f64 = _p(hold(tf.sh[sl])), _p(hold(tf.dbl[sl]))
        d.twi_u64, d.twi_f64 = _p(hold(ti.sh[sl])), _p(hold(ti.dbl[sl]))
        d.twpf_u64, d.twpf_f64 = _p(hold(tf.psh[sl])), _p(hold(tf.pdbl[sl]))
        d.twpi_u64, d.twpi_f64 = _p(hold(ti.psh[sl])), _p(hold(ti.pdbl[sl]))
        d.qinv = _p(hold(ntt.qinv[dev][sl]))
        d.sR, d.sR_sh = _p(T(ntt.fs_R[0])), _p(T(ntt.fs_R[1]))
        d.sExit, d.sExit_sh = _p(T(ntt.fs_exit[0])), _p(T(ntt.fs_exit[1]))
        d.PiR = _p(hold(eng._moddown_table(level, dev)))
        d.part_alpha, d.Lenter = _p(part_alpha), _p(lenter_ptrs)
        d.loc_row0, d.loc_alpha, d.loc_Y, d.loc_Ltri = _p(loc_row0), _p(loc_alpha), _p(loc_Y), _p(loc_L)
        if level > 0 and dev < len(eng.rescale_scales[level - 1]):
            d.rescale_scale = _p(hold(eng.rescale_scales[level - 1][dev]))
            src = ntt.p.rescaler_loc[level - 1]
            d.round_at = eng.ctx.q[ntt.p.destination_arrays[level - 1][src][0]] // 2
        # Horner tables of the extension fused into the forward column pass (FP64 targets)
        qrows = [eng.ctx.q[i] for i in ntt.p.d_special[dev][a:b]]
        SMALL = 1 << 42
        f64 = lambda v: hold(torch.tensor(v, dtype=torch.float64, device=device))
        wide, hm_ptrs = [], []
        mixed = False
        for s_ in self.sids:
            src, part_id, alpha = owners[s_]
            primes = ntt.parts_pack[src][tuple(ntt.p.p[level][src][part_id])]["prime_ids"]
            m = [eng.ctx.q[i] for i in primes]
            is_wide = any(x >= SMALL for x in m)
            if is_wide and alpha != 1 and any(qt < SMALL for qt in qrows):
                # scale_bits = 42: the table's primes alternate around 2^42, so a multi-limb partition can hold digits
                # beyond 2^51 while some target limbs are FP64 limbs.  The FP64 extension splits only ONE wide digit
                # (the base-prime partition): such contexts take the integer key-switch pipeline (natural-order keys).
                mixed = True
            wide.append(1 if is_wide else 0)
            if alpha > 1:
                hm = f64([[float(m[i] % qt) for qt in qrows] for i in range(alpha - 1)])
                hm_ptrs.append(hm.data_ptr())
            else:
                hm_ptrs.append(0)
        d.part_wide = _p(i32(wide))
        d.Hm = _p(i64(hm_ptrs))
        d.Rd = _p(f64([float(eng.ctx.R % qt) for qt in qrows]))
        d.C31 = _p(f64([float((1 << 31) % qt) for qt in qrows]))
        d.Rinv = _p(f64([float(pow(eng.ctx.R, -1, qt)) for qt in qrows]))
        specials = eng.ctx.q[-K:][::-1]
        d.Pinv = _p(f64([[float(pow(Pj, -1, qt)) if qt != Pj else 0.0 for qt in qrows] for Pj in specials]))
        n_small = 0
        while n_small < self.L and qrows[n_small] < SMALL:
            n_small += 1
        d.L_small = n_small
        d.amax = max(2, max(owners[s_][2] for s_ in self.sids))
        self.fp64 = not mixed    # Hm / Pinv tables present: the FP64 slab pipeline (the integer fall-back needs natural-order keys)
        if mixed:
            d.Hm, d.Pinv = None, None
        self.desc = d
        self.ref = ctypes.byref(d)
        self._keep = keep

        # persistent buffers (constant addresses): tensor-stage scratch, digits, key-switch scratch
        ws = eng._workspace(dev)
        L = self.L
        self.x = ws.get("x", 4 * L * N).view(4, L, N)
        self.d = ws.get("d", 3 * L * N).view(3, L, N)
        self.digits = ws.get("digits", max(L, 1) * N).view(max(L, 1), N)
        self.ks_ws = ws.get("ks", int(lib.ckks_exec_keyswitch_ws_elems(L, K, len(self.sids), N)))
        self.peer = {}                 # sid -> persistent copy of a remote partition's digits on this device
        self._digit_ptrs = None
        self._key_ptrs = {}

    def local_state(self, sid):
        i = self.local_sids.index(sid)
        r0, al = self.loc_row0[i], self.loc_alpha[i]
        return self.digits[r0:r0 + al]

    def digit_pointer_table(self, blocks):
        """blocks: {sid: [alpha,N] tensor with CONSTANT address}; uploaded once"""
        if self._digit_ptrs is None:
            self._digit_ptrs = torch.tensor([blocks[s].data_ptr() for s in self.sids], dtype=torch.int64,
                                            device=self.x.device)
            self._blocks = blocks
        return self._digit_ptrs

    def permuted_keys(self):
        """does this plan's key-switch stage keep its NTT domain in warp-interleaved order (option 18, FP64 pipeline only)?"""
        return bool(lib.ckks_get_option(18)) and self.fp64

    def key_pointer_tables(self, eng, ksk, hoist_g=0):
        """device tables of the row-0 pointers of every partition's key halves at this level -> (k0, k1, stride, permuted).
        With option 18 (default) the stage keeps its NTT-domain data in warp-interleaved order, so the pointers go to
        PERMUTED copies of the key rows (made once per key and device by ckks_perm_rows, cached on the engine)."""
        ck = (id(ksk.data), hoist_g)
        hit = self._key_ptrs.get(ck)
        permuted = bool(lib.ckks_get_option(18)) and self.fp64
        if hit is None or hit[0] is not ksk.data or hit[4] != permuted:
            start = eng.ntt.starts[self.level][self.dev]
            p0, p1 = [], []
            for s in self.sids:
                src, part_id, _alpha = self.owners[s]
                kd = ksk.data[eng.parts_alloc[self.level][src][part_id]].data
                h0, h1 = kd[0][self.dev], kd[1][self.dev]
                if hoist_g:     # hoisted rotation: the NTT-domain pre-image of the key under the Galois map (made once per key)
                    h0, h1 = eng._hoist_key(h0, hoist_g, permuted), eng._hoist_key(h1, hoist_g, permuted)
                elif permuted:
                    h0, h1 = eng._permuted_key(h0), eng._permuted_key(h1)
                p0.append(h0[start:].data_ptr())
                p1.append(h1[start:].data_ptr())
            dev = self.x.device
            stride = self.N if (permuted or hoist_g) else ksk.data[0].data[0][self.dev].stride(0)
            hit = (ksk.data, torch.tensor(p0, dtype=torch.int64, device=dev), torch.tensor(p1, dtype=torch.int64, device=dev),
                   stride, permuted)
            self._key_ptrs[ck] = hit
        return hit[1], hit[2], hit[3], hit[4]


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def tensor_stage(plan, polys, r0s):
    """polys: 4 tensors [L,N] (rows surviving the rescale, common row stride); r0s: 4 tensors [N] on this device"""
    s = polys[0].stride(0)
    with torch.cuda.device(plan.x.device):      # the C side launches on (and takes its side streams from) the CURRENT device
        check(lib.ckks_exec_tensor_stage(plan.ref, *[_p(t) for t in polys], s, *[_p(t) for t in r0s], _p(plan.x), _p(plan.d),
                                         _p(plan.digits), _stream(plan.x)), "exec_tensor_stage")


def digits_stage(plan, a, galois=0):
    """galois != 0: digits of the Galois image of `a` (canonical rows), gathered from the unrotated polynomial"""
    with torch.cuda.device(a.device):
        check(lib.ckks_exec_digits(plan.ref, _p(a), a.stride(0), _p(plan.digits), plan.N, int(galois), _stream(a)), "exec_digits")


def keyswitch_stage(plan, digit_ptrs, k0p, k1p, kstride, permuted, add0, add1, out0, out1, add0_galois=0, phase=3, parts=(-1, -1)):
    """add0_galois != 0: the addend of output 0 is the Galois image of add0, gathered inside the ModDown kernel.
    phase 1: extend + NTT only (k*/add*/out* may be None); phase 2: inner product + tail on the block phase 1 left behind"""
    add = add0 if add0 is not None else add1
    dev = plan.x.device
    with torch.cuda.device(dev):
        check(lib.ckks_exec_keyswitch_stage(plan.ref, _p(digit_ptrs), plan.N, _p(k0p), _p(k1p), kstride or 0, 1 if permuted else 0,
                                            _p(add0), _p(add1), add.stride(0) if add is not None else 0, int(add0_galois),
                                            _p(out0), _p(out1), plan.N, _p(plan.ks_ws), int(phase), int(parts[0]), int(parts[1]),
                                            _stream(plan.x)),
              "exec_keyswitch_stage")
