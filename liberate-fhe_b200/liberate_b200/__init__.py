"""liberate_b200 -- B200-native (sm_100a) RNS-CKKS mult/rotate hot path behind the
Desilo/liberate-fhe ``liberate.fhe.ckks_engine`` / ``liberate.ntt.ntt_cuda`` interfaces.

Importing the package loads libckks_b200.so; it fails loudly if the CUDA library is missing
(there is no CPU fallback).  ``install_as_liberate()`` aliases the package as ``liberate`` so that
code written against the reference (``from liberate import fhe``) runs unchanged.
"""
from . import _lib  # noqa: F401  (loads the CUDA library or raises)
from . import ntt  # noqa: F401

__version__ = "0.1.0"


def install_as_liberate():
    import sys
    pkg = sys.modules[__name__]
    sys.modules.setdefault("liberate", pkg)
    for name, mod in list(sys.modules.items()):
        if name.startswith(__name__ + "."):
            sys.modules.setdefault("liberate" + name[len(__name__):], mod)
    return pkg
