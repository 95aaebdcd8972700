"""Typed exceptions of the engine API (same class names / constructor arguments as the reference's
src/liberate/fhe/presets/errors.py, so ``except errors.MaximumLevelError`` keeps working) and the
``log_error`` decorator that logs and re-raises."""
import functools
import logging


def log_error(func):
    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        try:
            return func(*args, **kwargs)
        except Exception as e:
            logging.error(f"[Error] Error in {func.__name__} : {e}")
            raise
    return wrapper


class _EngineError(Exception):
    def __init__(self, message):
        self.message_error = message
        super().__init__(message)

    def __str__(self):
        return self.message_error

    __repr__ = __str__


class TestException(_EngineError):
    __test__ = False

    def __init__(self):
        super().__init__("test exception")


class NotFoundMessageSpecialPrimes(_EngineError):
    def __init__(self, message_bit, N):
        super().__init__(f"Can't find message_bit = {message_bit} and N = {N}")


class NotFoundScalePrimes(_EngineError):
    def __init__(self, scale_bits, N):
        super().__init__(f"Can't find scale bits = {scale_bits} and N = {N}")


class NotEnoughPrimes(_EngineError):
    def __init__(self, scale_bits, N):
        super().__init__(f"Not enough scale primes at scale bits = {scale_bits} and N = {N}")


class ViolatedAllowedQbits(_EngineError):
    def __init__(self, scale_bits, N, num_scales, max_qbits, total_qbits):
        super().__init__(f"Maximum allowed qbits are violated: scale_bits={scale_bits}, N={N}, "
                         f"num_scales={num_scales}: max_qbits={max_qbits} < requested total {total_qbits}")


class NotEnoughPrimesForBiasGuard(_EngineError):
    def __init__(self, bias_guard, num_special_primes):
        super().__init__(f"Guarding against biased overflow requires more than 2 special prime channels "
                         f"(bias_guard={bias_guard}, num_special_primes={num_special_primes})")


class NotFindBufferBitLength(_EngineError):
    def __init__(self, buffer_bit_length):
        super().__init__(f"Can't find buffer length bit {buffer_bit_length}; only 62 is supported by the sm_100a kernels")


class SecretKeyNotIncludeSpecialPrime(_EngineError):
    def __init__(self):
        super().__init__("The input secret key must include special prime channels.")


class DifferentTypeError(_EngineError):
    def __init__(self, a, b):
        super().__init__(f"The data types are different: {a}, {b}")


class NotMatchType(_EngineError):
    def __init__(self, origin, to):
        super().__init__(f"The data_struct origin should be a '{to}', but it is '{origin}'.")


class NotMatchDataStructState(_EngineError):
    def __init__(self, origin: str):
        super().__init__(f"Wrong format of the source {origin} detected: apply ntt / the Montgomery transformation first.")


class MaximumLevelError(_EngineError):
    def __init__(self, level, level_max):
        super().__init__(f"The number of multiplications available is exhausted: level {level} of at most {level_max}.")


class DeviceSelectError(_EngineError):
    def __init__(self):
        super().__init__("Unable to select the requested devices.")
