"""where does create_evk spend its time?  per-operator CUDA-event timing of one create_public_key(include_special=True) at gold,
ours vs the reference engine (oracle/_ref/site)."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "liberate-fhe_b200"):
    sys.path.insert(0, str(p))
from liberate_b200 import fhe  # noqa: E402
from liberate_b200.fhe.presets import params  # noqa: E402

kw = {k: v for k, v in params["gold"].items() if k != "devices"}
eng = fhe.ckks_engine(devices=[0], **kw)
sk = eng.create_secret_key()


def timed(label, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label:46s} gpu {e0.elapsed_time(e1) / reps * 1e3:9.1f} us   wall {(time.perf_counter() - t0) / reps * 1e6:9.1f} us", flush=True)
    return out


mt = -2
e = timed("rng.discrete_gaussian(repeats=1)", lambda: eng.rng.discrete_gaussian(repeats=1))
et = timed("ntt.tile_unsigned(e)", lambda: eng.ntt.tile_unsigned(eng.rng.discrete_gaussian(repeats=1), 0, mt))
timed("ntt.enter_ntt(e)", lambda: eng.ntt.enter_ntt([x.clone() for x in et], 0, mt))
a = timed("rng.randint(q lists, repeats=K)", lambda: eng.rng.randint(eng._q_lists(0, mt), repeats=eng.ctx.num_special_primes))
sa = timed("ntt.mont_mult(a, sk)", lambda: eng.ntt.mont_mult(a, sk.data, 0, mt))
timed("ntt.mont_sub(e, sa)", lambda: eng.ntt.mont_sub(et, sa, 0, mt))
timed("create_public_key(include_special=True)", lambda: eng.create_public_key(sk, include_special=True))
timed("create_evk", lambda: eng.create_evk(sk), reps=5)
timed("create_rotation_key", lambda: eng.create_rotation_key(sk, 1), reps=5)
timed("encorypt", lambda: eng.encorypt(eng.example(-1, 1), eng.create_public_key(sk)), reps=5)
try:
    from oracle import ref_engine
    ref_fhe, cache = ref_engine.load()
    ref = ref_fhe.ckks_engine(devices=[0], cache_folder=cache, **kw)
    rsk = ref.create_secret_key()
    timed("REF rng.discrete_gaussian(repeats=1)", lambda: ref.rng.discrete_gaussian(repeats=1))
    timed("REF rng.randint", lambda: ref.rng.randint(ref.ctx.q if False else [ref.ntt.qlists[0]], repeats=ref.ctx.num_special_primes))
    timed("REF create_public_key(include_special=True)", lambda: ref.create_public_key(rsk, include_special=True))
    timed("REF create_evk", lambda: ref.create_evk(rsk), reps=5)
    timed("REF create_rotation_key", lambda: ref.create_rotation_key(rsk, 1), reps=5)
except Exception as ex:  # the comparison rows are optional
    print("reference rows skipped:", repr(ex)[:200])
