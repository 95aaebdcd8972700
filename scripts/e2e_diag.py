"""where does the e2e step go?  times H2D of the operands, the mult, and D2H of the product separately"""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "liberate-fhe_b200")); sys.path.insert(0, str(ROOT))
from liberate_b200 import fhe
from liberate_b200.fhe.presets import params
eng = fhe.ckks_engine(devices=[0], **{k: v for k, v in params["gold"].items() if k != "devices"})
sk = eng.create_secret_key(); pk = eng.create_public_key(sk); evk = eng.create_evk(sk)
m = np.random.default_rng(1).uniform(-1, 1, eng.num_slots) + 0j
a, b = eng.encorypt(m, pk), eng.encorypt(m, pk)
dev = torch.device("cuda:0")
ha = [[t.cpu().pin_memory() for t in poly] for poly in a.data]
hb = [[t.cpu().pin_memory() for t in poly] for poly in b.data]
print("pinned:", ha[0][0].is_pinned(), ha[0][0].shape, ha[0][0].is_contiguous())
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e = [ev() for _ in range(4)]
    t0 = time.perf_counter()
    e[0].record()
    da = [[t.to(dev, non_blocking=True) for t in poly] for poly in ha]
    db = [[t.to(dev, non_blocking=True) for t in poly] for poly in hb]
    e[1].record()
    t1 = time.perf_counter()
    r = eng.mult(a._replace(data=da), b._replace(data=db), evk)
    e[2].record()
    t2 = time.perf_counter()
    outs = [[torch.empty_like(t, device="cpu").pin_memory() for t in poly] for poly in r.data] if rep == 0 else outs
    t3 = time.perf_counter()
    for poly, hp in zip(r.data, outs):
        for t, h in zip(poly, hp):
            h.copy_(t, non_blocking=True)
    e[3].record()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"rep {rep}: GPU ms h2d {e[0].elapsed_time(e[1]):.2f} mult {e[1].elapsed_time(e[2]):.2f} d2h {e[2].elapsed_time(e[3]):.2f} | "
          f"CPU ms issue-h2d {1e3*(t1-t0):.2f} issue-mult {1e3*(t2-t1):.2f} alloc {1e3*(t3-t2):.2f} issue-d2h+sync {1e3*(t4-t3):.2f}")
# resident mult CPU issue time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): eng.mult(a, b, evk)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"resident: CPU issue {1e2*(t1-t0):.3f} ms/mult, total {1e2*(t2-t0):.3f} ms/mult")
