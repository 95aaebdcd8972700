"""One-process-per-GPU (torch.distributed / NCCL) run of the golden engine flow: every rank owns logical device
`rank` of the reference's limb partition and must reproduce that device's golden digests bit for bit.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_flow_check.py
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "tests", ROOT / "liberate-fhe_b200"):
    sys.path.insert(0, str(p))

import flows  # noqa: E402
from golden_utils import Checker  # noqa: E402
from seeded_rng import SeededCsprng  # noqa: E402


def own_sampler_chain(fhe, params, rank, world, local):
    """the engine's OWN sampler (key and nonce broadcast from rank 0, ckks_engine._shared_key_material), before and after
    refresh(): keys and ciphertexts made by the ranks together must decrypt after a chain that makes every device the
    source of a rescale.  (Ranks that disagree on the repeated channels produce ciphertexts that still decrypt at the
    levels where only device 0's limbs are read -- the golden flow, with its injected streams, cannot see that.)"""
    failures = []
    eng = fhe.ckks_engine(devices=[f"cuda:{local}"] * world, distributed=True, fast=True, **params)
    rs = np.random.default_rng(3)
    m = rs.uniform(-1, 1, eng.num_slots) + 1j * rs.uniform(-1, 1, eng.num_slots)
    m /= np.abs(m).max() * 1.5
    w = 0.5 * np.exp(0.3j)
    for label in ("fresh", "after refresh()"):
        sk = eng.create_secret_key()
        pk = eng.create_public_key(sk)
        evk = eng.create_evk(sk)
        rotk = eng.create_rotation_key(sk, 1)
        x = eng.encorypt(m, pk)
        cw = eng.encorypt(np.full(eng.num_slots, w), pk)
        v, sources = m, set()
        while x.level < eng.num_levels - 1:
            sources.add(eng.ntt.p.rescaler_loc[x.level])
            x = eng.rotate_single(eng.mult(eng.add(x, x), cw, evk), rotk)
            v = np.roll(2 * v * w, 1)
        out = eng.decrode(x, sk)
        if world > 1 and len(sources) < 2:
            failures.append(f"[own sampler, {label}] the chain never rescaled from a device other than {sources}")
        if rank == 0:
            err = float(np.abs(out - v).max())
            if not err < 1e-3:
                failures.append(f"[own sampler, {label}] chain to level {x.level} decrypts with error {err}")
        eng.refresh()
    return failures


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from liberate_b200 import fhe
    G = ROOT / "tests" / "golden"
    g = json.loads((G / f"engine_D{world}.json").read_text())
    full = np.load(G / f"engine_D{world}_full.npz")
    failures = []
    for mode in ("executor", "fast-python"):
        eng = fhe.ckks_engine(devices=[f"cuda:{local}"] * world, distributed=True, fast=True, **g["params"])
        eng.use_executor = mode == "executor"
        eng.rng = SeededCsprng(eng.ctx.N, [len(d) for d in eng.ntt.p.d], max(eng.ntt.num_special_primes, 2),
                               devices=eng.ntt.devices, only_device=None)
        # every rank draws the full per-device lists (keeps the streams in lock step) and keeps its own entry
        real = eng.rng

        class Mine:
            def __getattr__(self, name):
                fn = getattr(real, name)
                if name == "randround":
                    return fn

                def call(*a, **k):
                    out = fn(*a, **k)
                    return [t if d == rank else None for d, t in enumerate(out)]
                return call
        eng.rng = Mine()
        plain_encode = eng.encode

        def encode(m, level=0, padding=True, _enc=plain_encode):
            if rank != 0:   # rank 0 consumes N uniform draws in randround; keep the seeded streams aligned
                real.randround(torch.zeros(eng.ctx.N, dtype=torch.float64, device="cuda"))
            return _enc(m, level, padding)
        eng.encode = encode
        chk = Checker(g["digests"], full, eng.ntt.devices)
        objs = flows.hot_path_flow(eng, chk)
        flows.extra_flow(eng, chk, objs)
        failures += [f"[{mode}] {f}" for f in chk.failures]
        if mode == "executor":     # the wire format in distributed mode: every rank assembles the whole prime-ordered tensor
            host = eng.cpu(objs["ct_ab"])
            want = np.concatenate([full[f"ct_ab_host/{c}"] for c in (0, 1)]) if "ct_ab_host/0" in full else None
            back = eng.cuda(host)
            for c in (0, 1):
                mine, again = objs["ct_ab"].data[c][rank], back.data[c][rank]
                if mine is not None and not torch.equal(mine, again):
                    failures.append(f"[{mode}] cpu()/cuda() round trip differs on rank {rank}")
            del want
    failures += own_sampler_chain(fhe, g["params"], rank, world, local)
    ok = torch.tensor([0 if failures else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if failures:
        print(f"rank {rank}: {len(failures)} mismatches, first: {failures[:3]}")
    if rank == 0:
        print("DIST_FLOW_OK" if int(ok.item()) == 1 else "DIST_FLOW_FAILED", f"world={world}")
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
