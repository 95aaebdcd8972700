"""Exports the reference's cached prime TABLES (data, not code) and its security-table lookups into
liberate_b200/fhe/cache/primes.json.  Bit-exact ciphertexts require the very same RNS primes, which the
reference ships as pickles (src/liberate/fhe/cache/resources/*.pkl) and selects with maximum_qbits()
(src/liberate/fhe/context/security_parameters.py).  Run in the build container only:

    python scripts/export_prime_tables.py
"""
import json
import pickle
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/src/liberate/fhe")
out = {}
res = REF / "cache/resources"
msp = pickle.load(open(res / "message_special_primes.pkl", "rb"))
sp = pickle.load(open(res / "scale_primes.pkl", "rb"))
out["message_special_primes"] = {str(bits): {str(N): [int(x) for x in v] for N, v in d.items()} for bits, d in msp.items()}
# entries the reference could not fill hold an error string (e.g. scale_bits 20/21 at N=2^17): kept as null
out["scale_primes"] = {f"{bits},{N}": ([int(x) for x in v] if isinstance(v, list) else None) for (bits, N), v in sp.items()}

# maximum_qbits(N, security_bits, quantum, distribution) for every table entry + the ring sizes we support
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
import importlib.util
spec = importlib.util.spec_from_file_location("secpar", REF / "context/security_parameters.py")
secpar = importlib.util.module_from_spec(spec)
spec.loader.exec_module(secpar)
mq = {}
for sec in (128, 192, 256):
    for quantum in ("pre_quantum", "post_quantum"):
        for dist in ("uniform", "error", "ternary"):
            for logN in range(10, 18):
                try:
                    mq[f"{sec},{quantum},{dist},{logN}"] = float(secpar.maximum_qbits(2 ** logN, sec, quantum, dist))
                except Exception as e:  # noqa: BLE001
                    mq[f"{sec},{quantum},{dist},{logN}"] = None
out["maximum_qbits"] = mq
dst = ROOT / "liberate-fhe_b200/liberate_b200/fhe/cache/primes.json"
dst.write_text(json.dumps(out, separators=(",", ":")))
print("wrote", dst, dst.stat().st_size, "bytes")
