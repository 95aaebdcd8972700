from .encdec import decode, encode, rotate, conjugate  # noqa: F401
