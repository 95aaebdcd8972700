"""One gold mult with the limbs dealt to D logical devices that all live on ONE GPU (LocalComm): the kernels every rank of a
D-GPU run launches, one device after the other, for an ncu launch list (per-kernel durations of a 1/D shard).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
        python scripts/shard_launches.py 8
"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "liberate-fhe_b200"))
import numpy as np, torch
import bench
from liberate_b200 import fhe
D = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = fhe.ckks_engine(devices=[0] * D, **bench.preset_params("gold"))
sk = eng.create_secret_key(); pk = eng.create_public_key(sk); evk = eng.create_evk(sk)
m = np.random.default_rng(1).uniform(-1, 1, eng.num_slots)
a, b = eng.encorypt(m, pk), eng.encorypt(m, pk)
for _ in range(3):
    eng.mult(a, b, evk)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.mult(a, b, evk)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("rows per device:", [len(d) for d in eng.ntt.p.d], "partitions:", len(eng._part_owners(1)))
if len(sys.argv) > 2 and sys.argv[2] == "sweep":
    # sum over the D shards of one mult (every rank's kernels, one after the other, on this GPU) under the executor options
    from liberate_b200._lib import lib
    def timed(n=30):
        for _ in range(3):
            eng.mult(a, b, evk)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            eng.mult(a, b, evk)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    print("default", round(timed(), 4), "ms for all", D, "shards")
    for key, vals in ((22, (1, 2, 3)), (16, (0, 1)), (10, (1, 2, 3)), (11, (4, 24, 4000)), (9, (25, 100, 4000)), (2, (0, 1)), (19, (0, 1))):
        old = lib.ckks_get_option(key)
        for v in vals:
            lib.ckks_set_option(key, v)
            eng._plans.clear()
            print("option", key, "=", v, round(timed(), 4), "ms", flush=True)
        lib.ckks_set_option(key, old)
