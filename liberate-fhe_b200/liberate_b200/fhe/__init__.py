from . import context, encdec  # noqa: F401
from .ckks_engine import ckks_engine  # noqa: F401
from .data_struct import data_struct  # noqa: F401
from .presets import params  # noqa: F401
