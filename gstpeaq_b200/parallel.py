"""Multi-GPU plumbing: one process per GPU, pairs sharded contiguously, one
gather of the per-pair results.  The (ref,test) pairs are
independent -- all of the reference's state is per element instance
(gstpeaq.c:110-139) -- so there is no data-path collective; torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests) only carries the
final <= 128 B per pair."""
import numpy as np

from . import RESULT_DTYPE

__all__ = ["shard_range", "gather_results", "RESULT_DTYPE"]


def shard_range(n_pairs, rank, world_size):
    """(first, count) of the contiguous block of pairs owned by `rank`"""
    base, rem = divmod(n_pairs, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_results(local, n_pairs, device):
    """all_gather of the per-pair result rows; every rank returns the full
    array [n_pairs] in pair order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    rank = dist.get_rank()
    base, rem = divmod(n_pairs, world)
    max_count = base + (1 if rem else 0)
    row = RESULT_DTYPE.itemsize
    buf = np.zeros(max_count * row, dtype=np.uint8)
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    buf[:raw.size] = raw
    send = torch.from_numpy(buf).to(device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    out = np.zeros(n_pairs, dtype=RESULT_DTYPE)
    for r in range(world):
        first, count = shard_range(n_pairs, r, world)
        if count:
            rows = recv[r].cpu().numpy()[:count * row].view(RESULT_DTYPE)
            out[first:first + count] = rows
    return out
