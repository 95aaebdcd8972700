"""
csprng -- STAND-IN sampler with the call surface of the reference's ``liberate.csprng.Csprng``
(src/liberate/csprng/csprng.py:18-323): randint / discrete_gaussian / randround / refresh.

OUT OF SCOPE for this round: the reference's fused ChaCha20 + uniform /
CDT-discrete-Gaussian CUDA kernels are key-generation / encryption only and never run on the mult/rotate
path.  This class draws from torch's Philox generator ON THE GPU (no host round trip) so that the engine is
usable end to end; it is NOT a cryptographically secure generator and says so here.  Tests drive both
engines with tests/seeded_rng.SeededCsprng instead (the reference's generator cannot be seeded).

Shapes follow the reference: lists indexed by logical device; "repeated" channels are identical on every
device (same seed), non-repeated channels are independent per device.
"""
import os

import torch


class Csprng:
    def __init__(self, num_coefs=2 ** 15, num_channels=[8], num_repeating_channels=2, sigma=3.2, devices=None,
                 seed=None, nonce=None, local_ids=None):
        self.num_coefs = num_coefs
        self.devices = list(devices)
        self.num_devices = len(self.devices)
        self.local_ids = list(range(self.num_devices)) if local_ids is None else list(local_ids)
        self.shares = (list(num_channels) if len(num_channels) == self.num_devices
                       else [num_channels[0]] * self.num_devices)
        self.num_repeating_channels = num_repeating_channels
        self.sigma = sigma
        self.refresh(seed)

    def refresh(self, seed=None, nonce=None):
        base = int.from_bytes(os.urandom(7), "little") if seed is None else int(seed)
        self._rep = {d: torch.Generator(device=self.devices[d]).manual_seed(base) for d in self.local_ids}
        self._own = {d: torch.Generator(device=self.devices[d]).manual_seed(base + 1 + d) for d in self.local_ids}

    def _uniform(self, gen, device, bound):
        return torch.randint(0, int(bound), (self.num_coefs,), dtype=torch.int64, device=device, generator=gen)

    def randint(self, amax=3, shift=0, repeats=0):
        if not isinstance(amax, (list, tuple)):
            amax = [[amax] for _ in self.shares]
        out = []
        for d, am in enumerate(amax):
            if d not in self.local_ids:
                out.append(None)
                continue
            dev = self.devices[d]
            n_non = len(am) - repeats
            rows = [self._uniform(self._own[d], dev, am[i]) for i in range(n_non)]
            rows += [self._uniform(self._rep[d], dev, am[n_non + i]) for i in range(repeats)]
            out.append(torch.stack(rows) + shift)
        return out

    def discrete_gaussian(self, non_repeats=0, repeats=1):
        shares = non_repeats if isinstance(non_repeats, (list, tuple)) else [non_repeats] * self.num_devices
        out = []
        for d in range(self.num_devices):
            if d not in self.local_ids:
                out.append(None)
                continue
            dev = self.devices[d]
            non = torch.randn((shares[d], self.num_coefs), dtype=torch.float64, device=dev, generator=self._own[d])
            rep = torch.randn((repeats, self.num_coefs), dtype=torch.float64, device=dev, generator=self._rep[d])
            out.append(torch.round(torch.cat([non, rep], 0) * self.sigma).to(torch.int64))
        return out

    def randround(self, coef):
        """sign(x) * (floor|x| + Bernoulli(frac|x|)) as int64 (csprng/randround_cuda_kernel.cu:8-37)"""
        d = self.local_ids[0]
        ab = coef.abs()
        fl = torch.floor(ab)
        u = torch.rand(coef.shape, dtype=torch.float64, device=coef.device, generator=self._own[d])
        r = (fl + (u < (ab - fl))).to(torch.int64)
        return torch.where(coef < 0, -r, r)
