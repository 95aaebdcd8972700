"""
params.py -- TEST / BENCH INFRASTRUCTURE (like everything under oracle/): the prime chain of an engine preset, read
straight from the data file the product ships (liberate-fhe_b200/liberate_b200/fhe/cache/primes.json = the reference's own
pickled prime tables, src/liberate/fhe/cache/resources/*.pkl) WITHOUT importing the product package, so that bench.py's
reference arm loads no product code at all.  Restates the selection rule of ckks_context.__init__
(src/liberate/fhe/context/ckks_context.py:209-262): as many scale primes as the security budget allows, then the base
prime and the special primes.
"""
import json
import math
from pathlib import Path

PRIMES = Path(__file__).resolve().parents[1] / "liberate-fhe_b200" / "liberate_b200" / "fhe" / "cache" / "primes.json"
# presets/params.py:1-30
PRESETS = {"bronze": (14, 1), "silver": (15, 2), "gold": (16, 4), "platinum": (17, 6)}


def preset_chain(name, scale_bits=40, security_bits=128, quantum="post_quantum", distribution="uniform"):
    """-> (q list: scale primes, base prime, special primes; logN; number of special primes)"""
    logN, K = PRESETS[name]
    N = 1 << logN
    t = json.loads(PRIMES.read_text())
    message_special = t["message_special_primes"]["60"][str(N)]
    scale_primes = t["scale_primes"][f"{scale_bits},{N}"]
    max_qbits = int(t["maximum_qbits"][f"{security_bits},{quantum},{distribution},{logN}"])
    base_special = message_special[:1 + K]
    budget = max_qbits - sum(math.log2(p) for p in base_special)
    num_scales = 0
    budget -= math.log2(scale_primes[num_scales])
    while budget > 0:
        num_scales += 1
        budget -= math.log2(scale_primes[num_scales])
    return list(scale_primes[:num_scales]) + list(base_special), logN, K
