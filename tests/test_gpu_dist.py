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
