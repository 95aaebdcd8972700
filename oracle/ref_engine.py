"""
ref_engine.py -- loader for the reference package installed under oracle/_ref/site by
``python oracle/build_ref.py --engine`` (build container).  TEST / BENCH INFRASTRUCTURE ONLY: it gives the GPU
box the reference's own ckks_engine + CUDA kernels so that (a) our engine's outputs can be compared bit for bit
with the reference engine's on identical keys and ciphertexts, and (b) the reference's own mult+relin can be timed
on the same B200 (bench.py --impl reference_gpu).  Import-time accommodations only: a stub
``matplotlib`` (absent, imported at module scope by the reference), ``numpy.bool8`` (removed in numpy 2), and a
writable cache folder seeded with the reference's prime tables.
"""
import shutil
import sys
import tempfile
import types
from pathlib import Path

SITE = Path(__file__).resolve().parent / "_ref" / "site"


def available():
    return (SITE / "liberate" / ".complete").exists()


def load():
    """returns (fhe module of the reference, writable cache folder)"""
    import numpy as np
    if not available():
        raise RuntimeError("reference package not installed under oracle/_ref/site (python oracle/build_ref.py --engine)")
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if str(SITE) not in sys.path:
        sys.path.insert(0, str(SITE))
    import torch  # noqa: F401  (the extensions link against libtorch)
    from liberate import fhe
    cache = Path(tempfile.mkdtemp(prefix="refcache_"))
    for f in (SITE / "liberate/fhe/cache/resources").glob("*.pkl"):
        shutil.copy(f, cache / f.name)
    return fhe, str(cache)
