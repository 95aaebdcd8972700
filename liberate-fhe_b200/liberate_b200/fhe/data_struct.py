"""data_struct -- the container every key / ciphertext travels in (reference: src/liberate/fhe/data_struct.py:5-24).
Field names, order and defaults are part of the API contract (users pickle these and read the flags)."""
from typing import NamedTuple

from .version import VERSION


class data_struct(NamedTuple):
    data: tuple | list          # tensors: per polynomial -> per device [limbs_on_device, N] int64
    include_special: bool       # special-prime rows present
    ntt_state: bool             # NTT domain
    montgomery_state: bool      # Montgomery form
    origin: str                 # one of presets.types.origins
    level: int
    hash: str                   # sha256 of the parameter string + primes
    version: str = VERSION
