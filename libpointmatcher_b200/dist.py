"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed).

Two ways the path shards:
  pairs  — independent registrations, pair j -> rank j mod G, no data-path collective
  shard  — the queries of ONE registration split into contiguous column ranges, reference
           replicated; the select histograms and the normal-equation sums are all-reduced inside
           libpmgpu (comm.cu, NCCL), bootstrapped with an ncclUniqueId broadcast from rank 0.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous column range [lo, hi) of rank `rank`; ranges tile [0, n) exactly"""
    return (rank * n) // world, ((rank + 1) * n) // world


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
