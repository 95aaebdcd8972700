"""Host-side logic of the N>1 path on CPU: the two exchange steps of the engine (DistComm.gather_states = the ModUp
digit all_gather, DistComm.bcast = the rescale-limb broadcast) under torch.distributed with the gloo backend,
world_size 2 and 3, with uneven partition ownership exactly as rns_partition produces it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, L, K):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        from pathlib import Path
        root = Path(__file__).resolve().parents[1]
        sys.path.insert(0, str(root / "liberate-fhe_b200"))
        # import only the pure-python pieces (no CUDA library needed for this test)
        import importlib.util
        def load(name, rel):
            spec = importlib.util.spec_from_file_location(name, root / "liberate-fhe_b200" / "liberate_b200" / rel)
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            return m
        comm_mod = load("comm_mod", "fhe/comm.py")
        part_mod = load("part_mod", "ntt/rns_partition.py")
        p = part_mod.rns_partition(L, K, world)
        comm = comm_mod.DistComm(["cpu"] * world)
        N = 64
        for level in (0, 1, K + 1):
            # owners exactly as ckks_engine._part_owners builds them
            counts = [len(parts) for parts in p.p[level]]
            alloc = [a[-counts[d] - 1:-1] for d, a in enumerate(p.part_allocations)]
            lowest = min(min(a) for a in alloc if len(a) > 0)
            owners = {}
            for src in range(world):
                for part_id, part in enumerate(p.p[level][src]):
                    owners[alloc[src][part_id] - lowest] = (src, len(part))
            mk = lambda sid, alpha: (torch.arange(alpha * N, dtype=torch.int64).view(alpha, N) + 1000 * sid + 7 * level)
            local = {sid: mk(sid, alpha) for sid, (src, alpha) in owners.items() if src == rank}
            got = comm.gather_states(local, owners, [rank], N)
            assert set(got[rank]) == set(owners)
            for sid, (src, alpha) in owners.items():
                assert torch.equal(got[rank][sid], mk(sid, alpha)), (level, sid)
            # rescale-limb broadcast from the device holding the smallest live prime
            src = p.rescaler_loc[level]
            row = torch.arange(N, dtype=torch.int64) * (level + 3)
            out = comm.bcast(row if rank == src else None, src, range(world), shape=(N,))
            assert torch.equal(out[rank], row)
        # the sampler's key material: every rank ends up with rank 0's 8 + 2 words (and a second call gives new ones)
        seed, nonce = comm.shared_key_material()
        assert len(seed) == 8 and len(nonce) == 2 and all(0 <= w < 2 ** 32 for w in seed + nonce)
        mine = torch.tensor(seed + nonce, dtype=torch.int64)
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref), "ranks disagree on the sampler's key / nonce"
        again, _ = comm.shared_key_material()
        assert again != seed
        assert comm_mod.LocalComm(["cpu"]).shared_key_material() == (None, None)
        comm.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,L,K", [(2, 17, 2), (3, 9, 2), (2, 7, 3)])
def test_exchange_steps_under_gloo(world, L, K):
    mp.spawn(_worker, args=(world, _free_port(), L, K), nprocs=world, join=True)
