"""Multi-GPU plumbing of the hot path: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo
in the CPU tests).  The batch of trajectories / problems shards with NO data-path collective; the only collective of
the whole path is the all-reduce of the POD Gram matrix when the rows (DOFs) of the snapshot matrix are sharded."""
import os


def rank_world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_slice(total, rank, world):
    """Contiguous, balanced slice of `total` independent units for `rank` (first total % world ranks get one more)."""
    base, rem = divmod(int(total), int(world))
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def shard_rows(total_rows, rank, world, multiple=1):
    """Row range of a row-sharded snapshot matrix; boundaries rounded to `multiple` rows."""
    blocks = -(-int(total_rows) // multiple)
    s = shard_slice(blocks, rank, world)
    return slice(min(s.start * multiple, total_rows), min(s.stop * multiple, total_rows))


def allreduce_sum_(t, group=None):
    """In-place sum over ranks (the POD Gram all-reduce); no-op for a single process."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over ranks (device timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def gather_sharded(local, total, group=None):
    """All-gather equally-ordered contiguous shards (shard_slice layout) of a (local_count, ...) tensor back into a
    (total, ...) tensor on every rank.  Used by callers that want the whole batch's results on each rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return local
    world = dist.get_world_size(group)
    sizes = [shard_slice(total, r, world) for r in range(world)]
    pad = max(s.stop - s.start for s in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    return torch.cat([o[:s.stop - s.start] for o, s in zip(outs, sizes)], dim=0)
