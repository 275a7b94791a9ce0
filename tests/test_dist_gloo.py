"""world_size-2 gloo tests (CPU) of the host logic behind the N > 1 paths of bench.py."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libpointmatcher_b200 import dist as pd
    try:
        n = 1_000_003
        lo, hi = pd.shard_range(n, rank, world)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([hi - lo]))
        bounds = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(bounds, torch.tensor([lo, hi]))
        ok = int(sum(int(s) for s in sizes) == n and all(int(bounds[i][1]) == int(bounds[i + 1][0]) for i in range(world - 1)))
        ok &= int(int(bounds[0][0]) == 0 and int(bounds[-1][1]) == n)
        # timing rule: the multi-GPU time is the max over ranks
        ok &= int(pd.max_over_ranks(10.0 + rank) == 10.0 + world - 1)
        # the ncclUniqueId travels from rank 0 as opaque bytes
        uid = pd.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 0)
        ok &= int(uid == bytes(range(128)))
        # independent pairs: every pair exactly once
        mine = pd.pair_assignment(1024, rank, world)
        counts = torch.zeros(1024, dtype=torch.int64)
        counts[mine] = 1
        dist.all_reduce(counts)
        ok &= int(bool((counts == 1).all()))
        # sharded sums: partial normal equations add up to the unsharded ones (what comm.cu all-reduces)
        rng = np.random.default_rng(0)
        J = rng.normal(size=(1000, 6))
        part = torch.from_numpy(J[lo * 1000 // n: hi * 1000 // n].T @ J[lo * 1000 // n: hi * 1000 // n])
        dist.all_reduce(part)
        ok &= int(np.allclose(part.numpy(), J.T @ J))
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_world_size_2_host_logic():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: 1, 1: 1}


def test_shard_columns_tile_exactly():
    from libpointmatcher_b200 import dist as pd
    for n in (0, 1, 7, 1000, 1000003):
        for world in (1, 2, 3, 8):
            cols = np.concatenate([np.arange(n)[pd.shard_columns(n, k, world)] for k in range(world)])
            assert len(cols) == n and (np.sort(cols) == np.arange(n)).all()
            sizes = [len(np.arange(n)[pd.shard_columns(n, k, world)]) for k in range(world)]
            assert max(sizes) - min(sizes) <= pd.SHARD_CHUNK   # balanced to within one chunk


def test_shard_range_tiles_exactly():
    from libpointmatcher_b200 import dist as pd
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            r = [pd.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
