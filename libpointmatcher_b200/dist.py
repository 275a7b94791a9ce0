"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed).

Two ways the path shards:
  pairs  — independent registrations, pair j -> rank j mod G, no data-path collective
  shard  — the queries of ONE registration dealt out to the ranks in chunks of 4096 consecutive columns, round-robin (every
           rank gets pieces of every part of the scan, so the ranks' kNN times are balanced — with one contiguous range per
           rank, one rank's slice of a 2 M-point scan was 1.6 x more expensive than another's and the others waited for it
           inside the exchange — and every piece keeps the scan's full local density), reference replicated; the select histograms and the
           normal-equation sums are exchanged inside the kernels (comm.cuh), bootstrapped by init_comm below.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous column range [lo, hi) of rank `rank`; ranges tile [0, n) exactly"""
    return (rank * n) // world, ((rank + 1) * n) // world


SHARD_CHUNK = 4096


def shard_columns(n, rank, world, chunk=SHARD_CHUNK):
    """the columns of rank `rank`: chunks of `chunk` consecutive columns dealt out round-robin (chunk c -> rank c mod world), as an
    index array; the ranks' arrays tile [0, n) exactly.  Consecutive columns of a scan are neighbours in space, so a rank's subset
    keeps the full local density (the lanes of a warp stay as close together as on one GPU) while every rank gets pieces of
    every part of the scan (balanced kNN times)."""
    import numpy as np
    if world <= 1:
        return np.arange(n)
    idx = np.arange(n)
    return idx[(idx // chunk) % world == rank]


def shard_take(a, rank, world, chunk=SHARD_CHUNK):
    """a[shard_columns(len(a), rank, world)] as ONE strided copy (no index array: this sits inside a registration's timed region)"""
    import numpy as np
    a = np.asarray(a)
    n = len(a)
    if world <= 1:
        return a
    nchunks = n // chunk
    body = a[:nchunks * chunk].reshape((nchunks, chunk) + a.shape[1:])[rank::world]
    tail = a[nchunks * chunk:] if (nchunks % world == rank) else a[:0]
    out = np.empty((body.shape[0] * chunk + len(tail),) + a.shape[1:], a.dtype)
    out[:body.shape[0] * chunk].reshape(body.shape)[...] = body
    out[body.shape[0] * chunk:] = tail
    return out


def pair_assignment(n_pairs, rank, world):
    """static round-robin: pair j -> rank j mod world (evaluations/eval_solution.cpp:250-271 shards pairs over threads)"""
    return list(range(rank, n_pairs, world))


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (timing rule: a multi-GPU time is the slowest rank's)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_bytes(payload, src=0):
    """rank `src` passes bytes (e.g. the 128-byte ncclUniqueId), every rank returns them"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return payload
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def gather_bytes(payload):
    """every rank passes bytes, every rank returns the list of all ranks' bytes in rank order"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [payload]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, payload)
    return out


def init_comm(ctx, capi, peer=True, nccl=True):
    """shard `ctx` (capi.Context) over the default process group: peer mailboxes for the per-iteration exchanges
    (fused into the kernels, comm.cuh) and an NCCL communicator for the all-gather of sharded map normals"""
    rank, world = dist.get_rank(), dist.get_world_size()
    if nccl:
        uid = broadcast_bytes(capi.comm_unique_id() if rank == 0 else None, 0)
        ctx.comm_init(uid, rank, world)
    if peer:
        handles = gather_bytes(ctx.comm_peer_handle())
        ctx.comm_peer_init(handles, rank)
        dist.barrier()  # every mailbox is mapped everywhere before the first exchange
