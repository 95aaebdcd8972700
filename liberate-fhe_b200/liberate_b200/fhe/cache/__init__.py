"""prime tables shipped as data (primes.json, exported from the reference's pickles by
scripts/export_prime_tables.py; bit-exact ciphertexts need the very same primes)"""
import json
from functools import lru_cache
from pathlib import Path

path_cache = str(Path(__file__).resolve().parent)


@lru_cache(maxsize=1)
def tables():
    return json.loads((Path(path_cache) / "primes.json").read_text())
