"""Batched NTT throughput sweep (BASELINE.json configs[2]): N in 2^14..2^17 x L limbs, forward and inverse
separately, CUDA-event timed on torch's current stream.  Working set is rotated over > 2x L2 so every call
streams from HBM.  Writes gpurun_out/ntt_sweep.json.   python scripts/ntt_bench.py [--quick]"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "liberate-fhe_b200"))
sys.path.insert(0, str(ROOT))

from liberate_b200._lib import lib, check, option_defaults  # noqa: E402


def primes(logN, L):
    ctx = json.loads((ROOT / "tests/golden/context.json").read_text())["contexts"]
    q = [c for c in ctx if c["args"]["logN"] == 17][0]["q"]   # NTT-friendly for every N <= 2^17
    return (q * ((L + len(q) - 1) // len(q)))[:L]


def consts(q):
    R = 1 << 62
    M = (1 << 31) - 1
    k = [(R * pow(R, -1, x) - 1) // x for x in q]
    t = lambda v: torch.tensor(v, dtype=torch.int64, device="cuda")
    return dict(_2q=t([2 * x for x in q]), ql=t([x & M for x in q]), qh=t([x >> 31 for x in q]),
                kl=t([x & M for x in k]), kh=t([x >> 31 for x in k]), Rs=t([R * R % x for x in q]))


def run(logN, L, iters, pool_bytes=320 << 20):
    N = 1 << logN
    q = primes(logN, L)
    c = consts(q)
    g = torch.Generator(device="cuda").manual_seed(0)
    nbuf = max(2, int(pool_bytes // (L * N * 8)) + 1)
    bufs = [torch.randint(0, 1 << 40, (L, N), dtype=torch.int64, device="cuda", generator=g) for _ in range(nbuf)]
    # throughput does not depend on twiddle values; use a random table (correctness is covered by the tests)
    tw = torch.randint(0, 1 << 40, (L, N), dtype=torch.int64, device="cuda", generator=g)
    st = torch.cuda.current_stream().cuda_stream
    P = lambda t: t.data_ptr()

    def fwd(b):
        check(lib.ckks_ntt(P(b), N, L, logN, P(tw), N, None, P(c["_2q"]), P(c["ql"]), P(c["qh"]), P(c["kl"]), P(c["kh"]), st), "ntt")

    def inv(b):
        check(lib.ckks_intt(P(b), N, L, logN, P(tw), N, P(c["Rs"]), P(c["_2q"]), P(c["ql"]), P(c["qh"]), P(c["kl"]), P(c["kh"]), 2, st), "intt")

    # canonical-output fast transforms (FP64 butterflies for q < 2^42, Shoup for the 60-bit primes)
    twu = torch.randint(0, 1 << 40, (L, N, 2), dtype=torch.int64, device="cuda", generator=g)
    twd = torch.randint(0, 1 << 40, (L, N), dtype=torch.int64, device="cuda", generator=g).double()
    qd = torch.tensor(q, dtype=torch.int64, device="cuda")
    sc = c["Rs"]

    def ffwd(b, fi=0):
        check(lib.ckks_ntt_fast(P(b), N, L, L, logN, P(twu), P(twd), P(qd), None, None, fi, st), "ntt_fast")

    def finv(b, fi=0):
        check(lib.ckks_intt_fast(P(b), N, L, L, logN, P(twu), P(twd), P(qd), P(sc), P(sc), 0, fi, st), "intt_fast")

    out = {}
    variants = [("fwd", fwd, None), ("inv_exit_reduce", inv, None), ("fast_fwd", ffwd, None), ("fast_inv", finv, None),
                ("fast_fwd_int", lambda b: ffwd(b, 1), None), ("fast_inv_int", lambda b: finv(b, 1), None),
                ("fast_fwd_classic", ffwd, (3, 0)), ("fast_fwd_1stream", ffwd, (10, 1))]
    for name, fn, opt in variants:
        if opt:
            lib.ckks_set_option(*opt)
        for i in range(3):
            fn(bufs[i % nbuf])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(bufs[i % nbuf])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name] = dict(ms=ms, gbps=16.0 * L * N / (ms * 1e-3) / 1e9, limb_ntt_us=ms * 1e3 / L)
        for k, v in option_defaults().items():
            lib.ckks_set_option(k, v)
    return out


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    res = []
    for logN in ([16] if quick else [14, 15, 16, 17]):
        for L in ([36] if quick else [1, 4, 16, 36, 60]):
            r = run(logN, L, iters=20 if quick else 50)
            res.append(dict(logN=logN, L=L, **r))
            print(logN, L, {k: (round(v["gbps"], 1), round(v["limb_ntt_us"], 3)) for k, v in r.items()}, flush=True)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / ("ntt_sweep_quick.json" if quick else "ntt_sweep.json")).write_text(json.dumps(dict(when=time.time(), results=res), indent=1))
