"""Per-operation timings of the engine at the gold preset on one GPU (config 4's operations: mult, rescale is inside
mult, rotate; plus add / level_up / key generation / encrypt), ours next to the reference engine when it is installed
under oracle/_ref/site.  CUDA-event timed, resident operands.   python scripts/op_bench.py [--preset gold]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "liberate-fhe_b200"))
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser()
ap.add_argument("--preset", default="gold")
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def run(fhe, label, extra):
    params = {k: v for k, v in fhe.presets.params[args.preset].items() if k != "devices"}
    eng = fhe.ckks_engine(devices=[0], **params, **extra)
    # key generation, WARM: the first calls build per-level tables and make torch's caching allocator cudaMalloc ~400 MB per
    # evaluation key (8.7 ms instead of 2.3 ms for ours; whichever engine runs second in this process inherits the other's
    # pool, which is what made round 1's "15 ms vs 3.7 ms" row) -- so: three warm-up keys, each dropped before the next
    evk = None
    for _ in range(4):
        del evk
        torch.cuda.synchronize(); t0 = time.perf_counter(); sk = eng.create_secret_key(); pk = eng.create_public_key(sk); torch.cuda.synchronize(); t_keys = time.perf_counter() - t0
        t0 = time.perf_counter(); evk = eng.create_evk(sk); torch.cuda.synchronize(); t_evk = time.perf_counter() - t0
    rotk = eng.create_rotation_key(sk, 1)
    m = eng.example(-1, 1)
    a, b = eng.encorypt(m, pk), eng.encorypt(m, pk)
    prod = eng.mult(a, b, evk)
    out = {"impl": label, "preset": args.preset,
           "keygen_sk_pk_ms": round(t_keys * 1e3, 2), "keygen_evk_ms": round(t_evk * 1e3, 2),
           "encorypt_us": round(timed(lambda: eng.encorypt(m, pk), 10), 1),
           "mult_relin_us": round(timed(lambda: eng.mult(a, b, evk), args.iters), 1),
           "rotate_single_us": round(timed(lambda: eng.rotate_single(a, rotk), args.iters), 1),
           "rotate_single_level1_us": round(timed(lambda: eng.rotate_single(prod, rotk), args.iters), 1),
           "add_us": round(timed(lambda: eng.add(a, b), args.iters), 1),
           "rescale_us": round(timed(lambda: eng.rescale(a), args.iters), 1),
           "level_up_us": round(timed(lambda: eng.level_up(a, 3), args.iters), 1),
           "mult_scalar_us": round(timed(lambda: eng.mult(a, 0.5), args.iters), 1),
           "mc_mult_us": round(timed(lambda: eng.mult(m, a), 10), 1),
           "mc_add_us": round(timed(lambda: eng.add(m, a), 10), 1),
           "decrode_us": round(timed(lambda: eng.decrode(prod, sk), 10), 1)}
    err = float(np.abs(eng.decrode(eng.rotate_single(prod, rotk), sk) - np.roll(m * m, 1)).max())
    out["mult_rotate_decrypt_error"] = err
    if hasattr(eng, "capture"):
        g = eng.capture(eng.mult, a, b, evk)
        out["mult_relin_graph_us"] = round(timed(g.replay, args.iters), 1)
        g2 = eng.capture(eng.rotate_single, a, rotk)
        out["rotate_single_graph_us"] = round(timed(g2.replay, args.iters), 1)
    print(json.dumps(out), flush=True)


from liberate_b200 import fhe as ours  # noqa: E402
run(ours, "liberate_b200", {})
try:
    from oracle import ref_engine
    if ref_engine.available():
        ref_fhe, cache = ref_engine.load()
        run(ref_fhe, "reference (its own CUDA kernels, sm_100)", {"cache_folder": cache})
except Exception as e:  # the comparison row is optional
    print(json.dumps({"impl": "reference", "unavailable": repr(e)}))
