"""The C-ABI library loads (no GPU needed) and exports exactly the symbols include/ckks_b200.h declares,
with the argument counts the ctypes table uses.  No compute calls here."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def header_decls():
    text = (ROOT / "include" / "ckks_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|int64_t)\s+(ckks_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


def test_library_loads_and_exports_every_declared_symbol():
    from liberate_b200 import _lib
    decls = header_decls()
    assert len(decls) >= 20
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, n in decls.items():
        assert hasattr(_lib.lib, name), name
        assert len(_lib.SIGNATURES[name]) == n, name
    assert _lib.lib.ckks_abi_version() == _lib.ABI_VERSION == 4


def test_no_cpu_fallback():
    import torch
    from liberate_b200.ntt import ntt_cuda
    with pytest.raises(RuntimeError):
        ntt_cuda.reduce_2q([torch.zeros(2, 8, dtype=torch.int64)], [torch.zeros(2, dtype=torch.int64)])


def test_product_does_not_import_the_oracle():
    pkg = ROOT / "liberate-fhe_b200" / "liberate_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
