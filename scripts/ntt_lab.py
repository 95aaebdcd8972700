"""Kernel lab: times the fast forward / inverse transforms (whole call, column pass alone, block pass alone) for a
list of option sets on the key-switch shape ([parts*E rows, N], constants with period E) and checks that every
option set produces the same bits as the default kernels.
    python liberate-fhe_b200/csrc/build.py --lab          # libckks_b200_lab.so: adds the measurement knobs 5 and 21
    CKKS_B200_LIB=liberate-fhe_b200/csrc/libckks_b200_lab.so python scripts/ntt_lab.py [--logN 16] [--rows 380] [--period 38]
                                       [--big 5] [--perm 1] [--opts "2=28;11=400;21=1;21=2"]
knob 5 = skip the column (1) / block (2) pass, knob 21 = skip the butterflies (1) / the global loads and stores (2); with the
product library the per-pass columns are NaN and those knobs are refused."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "liberate-fhe_b200"))
sys.path.insert(0, str(ROOT))
from liberate_b200._lib import lib, check, option_defaults  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--logN", type=int, default=16)
ap.add_argument("--rows", type=int, default=380)
ap.add_argument("--period", type=int, default=38)
ap.add_argument("--big", type=int, default=5, help="60-bit limbs per period (integer path)")
ap.add_argument("--opts", default="")
ap.add_argument("--perm", type=int, default=1, help="NTT-domain side in warp-interleaved order (as the executor runs it)")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--nbuf", type=int, default=3, help="buffers rotated between calls (1 + small rows = L2-resident)")
args = ap.parse_args()

logN, rows, E = args.logN, args.rows, args.period
PERM = args.perm
N = 1 << logN
ctx = json.loads((ROOT / "tests/golden/context.json").read_text())["contexts"]
qall = [c for c in ctx if c["args"]["logN"] == 17][0]["q"]
small = [x for x in qall if x < (1 << 42)]
big = [x for x in qall if x >= (1 << 42)]
q = (small * 8)[:E - args.big] + (big * 8)[:args.big]
g = torch.Generator(device="cuda").manual_seed(0)
qd = torch.tensor(q, dtype=torch.int64, device="cuda")
# plain "twiddles": any residues < q give the same instruction stream; the bit-equality check only needs determinism
tw = (torch.randint(0, 1 << 62, (E, N), dtype=torch.int64, device="cuda", generator=g) % qd[:, None]).contiguous()
twu = torch.empty((E, N, 2), dtype=torch.int64, device="cuda")
twd = torch.empty((E, N), dtype=torch.float64, device="cuda")
twpu, twpd = torch.empty_like(twu), torch.empty_like(twd)
st = torch.cuda.current_stream().cuda_stream
P = lambda t: t.data_ptr()
check(lib.ckks_fast_tables(P(tw), P(qd), P(twu), P(twd), E, N, st), "tables")
check(lib.ckks_fast_pack(P(twu), P(twd), P(twpu), P(twpd), E, logN, st), "pack")
qinv = torch.tensor([1.0 / float(x) for x in q], dtype=torch.float64, device="cuda")
sc = (torch.randint(1, 1 << 62, (E,), dtype=torch.int64, device="cuda", generator=g) % qd).contiguous()
sc_sh = torch.tensor([((int(s) << 64) // int(m)) - (1 << 64) if ((int(s) << 64) // int(m)) >= (1 << 63) else (int(s) << 64) // int(m)
                      for s, m in zip(sc.tolist(), q)], dtype=torch.int64, device="cuda")
src = (torch.randint(0, 1 << 62, (rows, N), dtype=torch.int64, device="cuda", generator=g) % qd.repeat((rows + E - 1) // E)[:rows, None]).contiguous()
nbuf = args.nbuf
bufs = [src.clone() for _ in range(nbuf)]


def fwd(b):
    check(lib.ckks_ntt_fast(P(b), N, rows, E, logN, P(twu), P(twd), P(twpu), P(twpd), P(qd), P(qinv), None, None, 0, PERM, st), "ntt_fast")


def inv(b):
    check(lib.ckks_intt_fast(P(b), N, rows, E, logN, P(twu), P(twd), P(twpu), P(twpd), P(qd), P(qinv), P(sc), P(sc_sh), 0, 0, PERM, st), "intt_fast")


def set_opts(o):
    for k, v in option_defaults().items():
        lib.ckks_set_option(k, v)
    lib.ckks_set_option(5, 0)
    lib.ckks_set_option(21, 0)
    for k, v in o:
        lib.ckks_set_option(k, v)


def timeit(fn):
    for i in range(3):
        fn(bufs[i % nbuf])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.iters):
        fn(bufs[i % nbuf])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters * 1e3


optsets = [[]] + [[tuple(int(x) for x in kv.split("=")) for kv in o.split(",") if kv] for o in args.opts.split(";") if o]
ref = {}
for o in optsets:
    line = {}
    for name, fn in (("fwd", fwd), ("inv", inv)):
        set_opts(o)
        x = src.clone()
        fn(x)
        torch.cuda.synchronize()
        if not o:
            ref[name] = x
            same = True
        else:
            same = bool(torch.equal(x, ref[name]))
        set_opts(o)
        t_all = timeit(fn)
        has_skip = lib.ckks_get_option(5) >= 0          # lab builds only (-DCKKS_LAB)
        t_col = t_blk = float("nan")
        if has_skip:
            set_opts(o + [(5, 2)])
            t_col = timeit(fn)
            set_opts(o + [(5, 1)])
            t_blk = timeit(fn)
        line[name] = dict(us=round(t_all, 1), col=round(t_col, 1), blk=round(t_blk, 1), same=same,
                          gbps=round(16.0 * rows * N / t_all / 1e3, 1))
    print(json.dumps({"lib": Path(str(lib._cdll._name)).name, "opts": o, **line}), flush=True)
set_opts([])
