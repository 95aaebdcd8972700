"""
comm -- the only two cross-device data movements of the RNS-CKKS hot path, behind one interface:

  * bcast : one limb row (the rescale limb, engine.py:999-1011; the encoded plaintext, engine.py:327-331)
            from its owner device to the devices that need it;
  * gather_states : the ModUp digit blocks of every key-switch partition to every device
            (engine.py:778-810).

The reference moves both through pinned HOST buffers (GPU -> CPU -> GPU, one Python process driving all
GPUs).  Here:
  LocalComm : one process owns every logical device; tensors move GPU -> GPU directly (NVLink P2P when the
              devices differ, nothing at all when they are the same device).
  DistComm  : one process per GPU under torch.distributed (NCCL over NVLink/NVSwitch on the GPU box, gloo in
              the CPU tests).  Logical device id == rank.  The digit exchange is ONE all_gather of a padded
              [max_rows, N] block per rank; the rescale limb is ONE broadcast.  Special-prime limbs are
              replicated on every rank (part.py:36-37), so ModDown needs no communication at all.
"""
import torch


class LocalComm:
    def __init__(self, devices):
        self.devices = devices
        self.world = 1
        self.rank = 0
        self.local_ids = list(range(len(devices)))

    def bcast(self, t, src, dst_ids):
        """t lives on logical device `src`; returns {dst: tensor on dst}"""
        return {d: (t if self.devices[d] == str(t.device) or d == src else t.to(self.devices[d], non_blocking=True))
                for d in dst_ids}

    def gather_states(self, local_states, owners, dst_ids, N):
        """local_states: {sid: [alpha,N] tensor on its owner}; owners: {sid: (src_dev, alpha)} for ALL sids.
        returns {dst: {sid: tensor on dst}}"""
        out = {}
        for d in dst_ids:
            out[d] = {}
            for sid, (src, _alpha) in owners.items():
                t = local_states[sid]
                out[d][sid] = t if (d == src or self.devices[d] == str(t.device)) else t.to(self.devices[d], non_blocking=True)
        return out

    def barrier(self):
        pass

    def shared_key_material(self):
        """one process: the sampler draws its own key and nonce from os.urandom, as the reference (csprng.py:215-223)"""
        return None, None


class DistComm:
    """logical device id == torch.distributed rank; this process owns exactly one device"""

    def __init__(self, devices, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if len(devices) != self.world:
            raise ValueError(f"{len(devices)} logical devices for world size {self.world}")
        self.devices = devices
        self.local_ids = [self.rank]
        self.device = devices[self.rank]

    def shared_key_material(self):
        """collective: (seed, nonce) = 8 + 2 random 32-bit words drawn by rank 0 and handed to every rank.  All ranks must
        run the SAME ChaCha20 key and nonce, or the sampler's repeated channels (secret key, shared randomness of public keys
        and encryptions) differ between the ranks -- such ciphertexts decrypt only while device 0's limbs alone are read."""
        import os
        box = [None]
        if self.rank == 0:
            words = [int.from_bytes(os.urandom(4), "big") for _ in range(10)]
            box = [(words[:8], words[8:])]
        self.dist.broadcast_object_list(box, src=0, group=self.group)
        seed, nonce = box[0]
        return list(seed), list(nonce)

    def bcast(self, t, src, dst_ids, shape=None, dtype=torch.int64):
        """collective: every rank calls it; ranks other than `src` pass t=None (+ shape)"""
        if self.rank == src:
            buf = t.contiguous()
        else:
            buf = torch.empty(shape, dtype=dtype, device=self.device)
        self.dist.broadcast(buf, src=src, group=self.group)
        return {self.rank: buf} if self.rank in dst_ids else {}

    def gather_states(self, local_states, owners, dst_ids, N):
        """ONE all_gather: every rank contributes its digit rows (padded to the widest rank)"""
        rows_of = [0] * self.world
        for sid, (src, alpha) in owners.items():
            rows_of[src] += alpha
        width = max(rows_of)
        mine = torch.zeros((max(width, 1), N), dtype=torch.int64, device=self.device)
        r = 0
        for sid in sorted(local_states):
            t = local_states[sid]
            mine[r:r + t.size(0)] = t
            r += t.size(0)
        everything = torch.empty((self.world, max(width, 1), N), dtype=torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(everything.view(-1, N), mine, group=self.group)
        out = {self.rank: {}}
        cursor = [0] * self.world
        for sid in sorted(owners):
            src, alpha = owners[sid]
            out[self.rank][sid] = everything[src, cursor[src]:cursor[src] + alpha]
            cursor[src] += alpha
        return out if self.rank in dst_ids else {}

    def barrier(self):
        self.dist.barrier(group=self.group)
