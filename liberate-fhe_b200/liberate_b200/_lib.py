"""ctypes loader for libckks_b200.so (the C ABI declared in include/ckks_b200.h).

There is NO CPU fallback: if the CUDA library cannot be loaded the import fails loudly, and every
wrapper refuses tensors that are not on a CUDA device.
"""
import ctypes
import os
import shutil
import sys
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_CSRC = _PKG.parent / "csrc"
# CKKS_B200_LIB: load a differently-tuned build of the same sources (kernel experiments); default = the in-tree library
_LIBPATH = Path(os.environ["CKKS_B200_LIB"]).resolve() if os.environ.get("CKKS_B200_LIB") else _CSRC / "libckks_b200.so"

_i64p = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p

# name -> argtypes, mirroring include/ckks_b200.h exactly (checked by tests/test_abi.py)
SIGNATURES = {
    "ckks_abi_version": [],
    "ckks_set_option": [_int, _int],
    "ckks_get_option": [_int],
    "ckks_launch_count": [],
    "ckks_mont_mult": [_i64p, _i64, _i64p, _i64, _i64p, _i64, _int, _int, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_mont_enter": [_i64p, _i64, _i64p, _int, _int, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_ntt": [_i64p, _i64, _int, _int, _i64p, _i64, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_intt": [_i64p, _i64, _int, _int, _i64p, _i64, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p, _int, _vp],
    "ckks_mont_redc": [_i64p, _i64, _int, _int, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_reduce_2q": [_i64p, _i64, _int, _int, _i64p, _vp],
    "ckks_make_signed": [_i64p, _i64, _int, _int, _i64p, _vp],
    "ckks_make_unsigned": [_i64p, _i64, _int, _int, _i64p, _vp],
    "ckks_mont_add": [_i64p, _i64, _i64p, _i64, _i64p, _i64, _int, _int, _i64p, _vp],
    "ckks_mont_sub": [_i64p, _i64, _i64p, _i64, _i64p, _i64, _int, _int, _i64p, _vp],
    "ckks_addsub_reduce": [_i64p, _i64, _i64p, _i64, _i64p, _i64, _int, _int, _i64p, _int, _vp],
    "ckks_tile_unsigned": [_i64p, _i64p, _i64, _int, _int, _i64p, _vp],
    "ckks_compact_twiddles": [_i64p, _i64p, _int, _int, _int, _vp],
    "ckks_fast_tables": [_i64p, _i64p, _vp, _vp, _int, _int, _vp],
    "ckks_fast_pack": [_vp, _vp, _vp, _vp, _int, _int, _vp],
    "ckks_perm_rows": [_i64p, _i64, _i64p, _i64, _int, _int, _int, _vp],
    "ckks_ntt_fast": [_i64p, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _i64p, _vp, _i64p, _i64p, _int, _int, _vp],
    "ckks_intt_fast": [_i64p, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _i64p, _vp, _i64p, _i64p, _int, _int, _int, _vp],
    "ckks_rescale_scaled": [_i64p, _i64, _i64p, _i64p, _i64, _int, _int, _i64p, _i64p, _i64, _i64p, _i64p, _i64p, _i64p, _i64p,
                            _i64p, _vp],
    "ckks_pc_product": [_i64p, _i64p, _i64p, _i64, _i64p, _i64p, _i64, _int, _int, _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_pc_add": [_i64p, _i64, _i64p, _i64, _i64p, _i64, _int, _int, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_rescale": [_i64p, _i64, _i64p, _i64p, _i64, _int, _int, _i64p, _i64, _int, _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_tensor_product": [_i64p, _i64p, _i64p, _i64p, _i64, _i64p, _i64p, _i64p, _i64, _int, _int,
                            _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_garner_digits": [_i64p, _i64, _i64p, _i64, _int, _int, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_extend": [_i64p, _i64, _int, _i64p, _i64, _int, _int, _i64p, _i64p, _int, _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_ksk_inner": [_i64p, _i64, _int, _vp, _vp, _i64, _i64p, _i64p, _i64, _int, _int, _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_ksk_accumulate": [_i64p, _i64, _i64p, _i64p, _i64, _i64p, _i64p, _i64, _int, _int, _int,
                            _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_moddown": [_i64p, _i64, _int, _int, _int, _i64p, _i64p, _i64p, _i64, _i64p, _i64, _i64p,
                     _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_automorphism": [_i64p, _i64, _i64p, _i64, _int, _int, _i64, _int, _i64p, _vp],
    "ckks_exec_tensor_stage": [_vp, _i64p, _i64p, _i64p, _i64p, _i64, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p, _vp],
    "ckks_exec_digits": [_vp, _i64p, _i64, _i64p, _i64, _i64, _vp],
    "ckks_exec_keyswitch_stage": [_vp, _vp, _i64, _vp, _vp, _i64, _int, _i64p, _i64p, _i64, _i64, _i64p, _i64p, _i64, _i64p, _int, _int, _int, _vp],
    "ckks_exec_keyswitch_ws_elems": [_int, _int, _int, _int],
    "ckks_rng_bytes": [_i64p, _int, _int, _vp, _vp, _vp, ctypes.c_uint64, _vp],
    "ckks_rng_randint": [_i64p, _int, _int, _vp, _i64, _vp, _vp, _vp, ctypes.c_uint64, _vp],
    "ckks_rng_gaussian": [_i64p, _int, _int, _vp, _int, _int, _vp, _vp, _vp, ctypes.c_uint64, _vp],
    "ckks_rng_randround": [_vp, _i64p, _int, _vp, _vp, _vp, ctypes.c_uint64, _vp],
}
RESTYPES = {"ckks_exec_keyswitch_ws_elems": ctypes.c_int64, "ckks_launch_count": ctypes.c_int64}

ERRORS = {-1: "CKKS_E_BADARG (null pointer / bad size)", -2: "CKKS_E_LOGN (logN outside [12,17])",
          -3: "CKKS_E_ALIGN (pointer/stride not 16-byte aligned)"}


class CkksLibError(RuntimeError):
    pass


ABI_VERSION = 4      # the version SIGNATURES (and fhe/executor.LevelT) were written for


def _build_locked():
    """(re)build the in-tree library under a file lock: with one process per GPU every rank may find it missing or stale"""
    import fcntl
    sys.path.insert(0, str(_CSRC))
    try:
        import build as _b
        with open(_CSRC / ".build.lock", "w") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                _b.build()          # no-op when another rank has just built it
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    finally:
        sys.path.pop(0)


def _load():
    in_tree = not os.environ.get("CKKS_B200_LIB")
    have_nvcc = shutil.which("nvcc") is not None
    if not _LIBPATH.exists():
        if not (in_tree and have_nvcc):
            raise ImportError(f"{_LIBPATH} is missing and cannot be built here (nvcc: {have_nvcc}); "
                              "liberate_b200 has no CPU fallback")
        _build_locked()
    elif in_tree and have_nvcc:
        sys.path.insert(0, str(_CSRC))
        try:
            import build as _b
            stale = _b.stale()
        finally:
            sys.path.pop(0)
        if stale:
            _build_locked()
    lib = ctypes.CDLL(str(_LIBPATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = RESTYPES.get(name, ctypes.c_int)
    got = lib.ckks_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"{_LIBPATH} has ABI version {got}, this package was written for {ABI_VERSION}: "
                          "rebuild it (python liberate-fhe_b200/csrc/build.py --force)")
    return lib


class _Counted:
    """the loaded library; `launches` = kernels launched so far, counted inside the library itself
    (ckks_launch_count: one tick per kernel launch), so bench.py's gpu_launches is exact"""

    def __init__(self, cdll):
        self._cdll = cdll
        for name in SIGNATURES:
            setattr(self, name, getattr(cdll, name))

    @property
    def launches(self):
        return int(self._cdll.ckks_launch_count())


lib = _Counted(_load())
OPTION_KEYS = (2, 9, 10, 11, 12, 16, 17, 18, 19, 22)
_OPTION_DEFAULTS = {k: lib.ckks_get_option(k) for k in OPTION_KEYS}


def option_defaults():
    """the library's built-in knob values (before CKKS_B200_OPTIONS), e.g. to restore them after a test"""
    return dict(_OPTION_DEFAULTS)


# tuning switches of the library (ckks_set_option): CKKS_B200_OPTIONS="1=1,2=28"
for _kv in filter(None, os.environ.get("CKKS_B200_OPTIONS", "").split(",")):
    _k, _v = _kv.split("=")
    lib.ckks_set_option(int(_k), int(_v))


def check(rc, what):
    if rc != 0:
        msg = ERRORS.get(rc, f"cudaError {rc}")
        raise CkksLibError(f"{what} failed: {msg}")
