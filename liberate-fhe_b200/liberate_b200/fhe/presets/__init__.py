from .params import params  # noqa: F401
from . import types  # noqa: F401
from . import errors  # noqa: F401
