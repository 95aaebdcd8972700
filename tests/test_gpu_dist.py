"""One process per GPU (torch.distributed / NCCL): every rank owns logical device `rank` of the reference's limb partition
and must reproduce that device's golden digests bit for bit (tests/dist_flow_check.py: the whole golden flow through the
executor and through the Python-orchestrated fast path, both collectives included).  Needs >= 2 GPUs: skipped on a
one-GPU box (the CPU twin is tests/test_comm_gloo.py, world 2 and 3 over gloo)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("world", [2, 3])
def test_golden_flow_one_process_per_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    env = dict(os.environ)
    env.pop("CKKS_B200_OPTIONS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), str(ROOT / "tests" / "dist_flow_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=str(ROOT))
    assert out.returncode == 0 and "DIST_FLOW_OK" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])


def test_two_physical_devices_in_one_process():
    """the reference's main mode: ONE process, devices=[0, 1] (LocalComm, GPU->GPU copies).  The C executor launches on the
    CURRENT device and takes its side streams from it, so every executor call must set the tensors' device first (ADVICE r01:
    without the guard, logical device 1's kernels ran on device 0).  Golden flow for two logical devices on two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 GPUs, this box has {torch.cuda.device_count()}")
    import json

    import numpy as np

    import flows
    from conftest import GOLDEN
    from golden_utils import Checker
    from liberate_b200 import fhe
    from seeded_rng import SeededCsprng
    g = json.loads((GOLDEN / "engine_D2.json").read_text())
    full = np.load(GOLDEN / "engine_D2_full.npz")
    for mode in ("executor", "fast-python"):
        eng = fhe.ckks_engine(devices=[0, 1], fast=True, **g["params"])
        eng.use_executor = mode == "executor"
        eng.rng = SeededCsprng(eng.ctx.N, [len(d) for d in eng.ntt.p.d], max(eng.ntt.num_special_primes, 2), devices=eng.ntt.devices)
        assert eng.ntt.devices[0] != eng.ntt.devices[1]
        chk = Checker(g["digests"], full, eng.ntt.devices)
        objs = flows.hot_path_flow(eng, chk)
        flows.extra_flow(eng, chk, objs)
        for d in (0, 1):
            torch.cuda.synchronize(d)
        assert not chk.failures, (mode, chk.failures[:8])
        # hoisted rotations and the host round trip on two real devices
        outs = eng.rotate_hoisted(objs["ct_ab"], [objs["rotk1"]])
        want = np.roll(objs["ma"] * objs["mb"], 1)
        assert np.abs(eng.decrode(outs[0], objs["sk"]) - want).max() < 1e-6
        back = eng.cuda(eng.cpu(objs["ct_ab"]))
        assert all(torch.equal(x, y) for pa, pb in zip(back.data, objs["ct_ab"].data) for x, y in zip(pa, pb))
