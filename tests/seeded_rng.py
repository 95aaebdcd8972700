"""
seeded_rng.py -- deterministic stand-in for the reference's Csprng (src/liberate/csprng/csprng.py:18-323).

The reference generator cannot be seeded, so parity between the reference engine and
this repo's engine is established by driving BOTH with this sampler: tests/golden/make_golden.py
plugs it into the unmodified reference engine (build container), the tests plug it into
liberate_b200.  Same call surface, same shapes, same "repeated channels are identical on every
device" rule; the values are ordinary numpy draws.

``rank``/``world`` let one process of a sharded (one-process-per-GPU) engine reproduce exactly the
slice device ``rank`` would have received in the reference's single-process multi-device run.
"""
import numpy as np
import torch


class SeededCsprng:
    def __init__(self, num_coefs=2 ** 15, num_channels=[8], num_repeating_channels=2, sigma=3.2,
                 devices=None, seed=None, nonce=None, only_device=None):
        self.num_coefs = num_coefs
        self.devices = list(devices)
        self.num_devices = len(self.devices)
        self.shares = (list(num_channels) if len(num_channels) == self.num_devices
                       else [num_channels[0]] * self.num_devices)
        self.num_repeating_channels = num_repeating_channels
        self.sigma = sigma
        self.only_device = only_device
        self.rng = np.random.default_rng(12345 if seed is None else seed)

    def _out(self, per_device):
        """per_device: list of numpy arrays (one per logical device)"""
        if self.only_device is None:
            return [torch.from_numpy(a).to(d) for a, d in zip(per_device, self.devices)]
        return [torch.from_numpy(per_device[self.only_device]).to(self.devices[self.only_device])]

    def randint(self, amax=3, shift=0, repeats=0):
        if not isinstance(amax, (list, tuple)):
            amax = [[amax] for _ in self.shares]
        out = []
        rep = None
        for am in amax:
            n_non = len(am) - repeats
            rows = [self.rng.integers(0, int(am[i]), self.num_coefs, dtype=np.int64) for i in range(n_non)]
            if repeats:
                if rep is None:
                    rep = [self.rng.integers(0, int(am[n_non + i]), self.num_coefs, dtype=np.int64)
                           for i in range(repeats)]
                rows += rep
            out.append(np.stack(rows) + shift)
        return self._out(out)

    def discrete_gaussian(self, non_repeats=0, repeats=1):
        shares = non_repeats if isinstance(non_repeats, (list, tuple)) else [non_repeats] * self.num_devices
        rep = np.rint(self.rng.normal(0, self.sigma, (repeats, self.num_coefs))).astype(np.int64)
        out = []
        for dev in range(self.num_devices):
            non = np.rint(self.rng.normal(0, self.sigma, (shares[dev], self.num_coefs))).astype(np.int64)
            out.append(np.concatenate([non, rep], 0))
        return self._out(out)

    def randround(self, coef):
        # csprng/randround_cuda_kernel.cu:8-37: sign * (floor|x| + Bernoulli(frac)), returned as int64
        c = coef.detach().cpu().numpy()
        ab = np.abs(c)
        fl = np.floor(ab)
        r = (fl + (self.rng.random(c.shape) < (ab - fl))).astype(np.int64)
        return torch.from_numpy(np.where(np.signbit(c), -r, r)).to(coef.device)
