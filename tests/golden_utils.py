"""shared by tests/golden/make_golden.py (records) and tests/test_gpu_engine.py (checks)"""
import hashlib

import numpy as np
import torch


def sha(t):
    a = np.ascontiguousarray(t.detach().cpu().numpy())
    return hashlib.sha256(a.tobytes()).hexdigest()


def describe(obj):
    """nested description of a data_struct / list / tensor: shapes + sha256 of the int64 bytes"""
    if hasattr(obj, "_fields") and "montgomery_state" in obj._fields:      # data_struct (either package)
        return {"__ds__": dict(include_special=obj.include_special, ntt_state=obj.ntt_state,
                               montgomery_state=obj.montgomery_state, origin=obj.origin, level=obj.level),
                "data": describe(obj.data)}
    if isinstance(obj, (list, tuple)):
        return [describe(o) for o in obj]
    if isinstance(obj, torch.Tensor):
        return {"shape": list(obj.shape), "dtype": str(obj.dtype).replace("torch.", ""), "sha": sha(obj)}
    if isinstance(obj, np.ndarray):
        return {"shape": list(obj.shape), "dtype": str(obj.dtype),
                "sha": hashlib.sha256(np.ascontiguousarray(obj).tobytes()).hexdigest()}
    if obj is None:
        return None
    raise TypeError(type(obj))


class Recorder:
    def __init__(self):
        self.digests = {}
        self.full = {}

    def __call__(self, name, obj):
        self.digests[name] = describe(obj)

    def fix(self, name, obj):
        self(name, obj)
        for d, t in enumerate(obj):
            self.full[f"{name}/{d}"] = t.detach().cpu().numpy()
        return obj


def first_diff(got, want, path=""):
    """path of the first mismatch between two descriptions (None = ours is absent for a non-local device)"""
    if got is None:
        return None
    if isinstance(want, dict) and "__ds__" in want:
        if got.get("__ds__") != want["__ds__"]:
            return f"{path}.__ds__: {got.get('__ds__')} != {want['__ds__']}"
        return first_diff(got["data"], want["data"], path + ".data")
    if isinstance(want, list):
        if not isinstance(got, list) or len(got) != len(want):
            return f"{path}: length {len(got) if isinstance(got, list) else got} != {len(want)}"
        for i, (g, w) in enumerate(zip(got, want)):
            d = first_diff(g, w, f"{path}[{i}]")
            if d:
                return d
        return None
    if got != want:
        return f"{path}: {got} != {want}"
    return None


class Checker:
    """replays a golden recording: every rec(name, obj) must reproduce the recorded digests"""

    def __init__(self, digests, full, devices, tolerance_names=()):
        self.digests = digests
        self.full = full
        self.devices = devices
        self.seen = []
        self.failures = []
        self.tolerance_names = set(tolerance_names)

    def __call__(self, name, obj):
        self.seen.append(name)
        diff = first_diff(describe(obj), self.digests[name], name)
        if diff:
            self.failures.append(diff)

    def fix(self, name, obj):
        """FFT-dependent values (encode): must be within +-1 of the golden integers; the golden tensors are
        returned so that everything downstream sees identical inputs."""
        self.seen.append(name)
        out = []
        for d, t in enumerate(obj):
            want = torch.from_numpy(self.full[f"{name}/{d}"])
            if t is None:
                out.append(None)
                continue
            delta = (t.detach().cpu() - want).abs().max().item()
            if delta > 1:
                self.failures.append(f"{name}/{d}: encode differs from golden by {delta}")
            out.append(want.to(t.device))
        return out
