"""Generates tests/golden/csprng.npz from the REFERENCE's own Python code (build container only; needs
oracle/_ref/site, see oracle/build_ref.py --engine):
  * chacha20_naive.chacha20 (src/liberate/csprng/chacha20_naive.py:104-112) on seeded random states,
  * build_CDT_binary_search_tree (src/liberate/csprng/discrete_gaussian_sampler.py:12-116).
    python tests/golden/make_golden_csprng.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_engine  # noqa: E402

ref_engine.load()
import liberate.csprng.chacha20_naive as cn  # noqa: E402
import liberate.csprng.discrete_gaussian_sampler as dgs  # noqa: E402

rng = np.random.default_rng(20261017)
n = 64
states = rng.integers(0, 1 << 32, (16, n), dtype=np.int64)
states[0], states[1], states[2], states[3] = 1634760805, 857760878, 2036477234, 1797285236
states[13, : n // 2] = 0                      # typical: small counters
out = cn.chacha20(torch.from_numpy(states.copy())).numpy()
btree, _ptr, size, depth = dgs.build_CDT_binary_search_tree(security_bits=128, sigma=3.2)
flat = np.ascontiguousarray(btree.T.ravel(), dtype=np.uint64)
np.savez_compressed(ROOT / "tests/golden/csprng.npz", states=states, chacha20=out, tree=flat, tree_size=size, tree_depth=depth)
print("wrote", ROOT / "tests/golden/csprng.npz", out.shape, flat.shape, size, depth)
