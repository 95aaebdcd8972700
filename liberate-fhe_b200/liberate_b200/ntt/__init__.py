from . import ntt_cuda  # noqa: F401
