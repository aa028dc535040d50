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


def overlapped_gram_allreduce(X_local, nblocks, group=None, block_multiple=128):
    """G = sum over ranks of X_g^T X_g with the reduction of finished pieces overlapped with the contraction of the
    next ones: the upper triangle of G is cut into `nblocks` block rows; block row I is one DMMA GEMM
    X[:, I]^T X[:, I:] into its own contiguous buffer, and as soon as it is done a side stream all-reduces that buffer
    (NCCL, asynchronous) while the compute stream is already in block row I + 1.  Only the upper triangle crosses
    NVLink (ns^2 / 2 + diagonal blocks instead of ns^2 doubles); the lower one is mirrored locally at the end.
    Returns the full symmetric G (ns x ns) on every rank."""
    import torch
    import torch.distributed as dist
    from .mor.pod import dgemm_device
    from . import _lib as L
    nf, ns = X_local.shape
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    nblk = max(1, min(int(nblocks), -(-ns // block_multiple)))
    # equal-AREA block rows of the triangle: row I of height h covers h * (ns - c0) entries
    edges = [0]
    area = ns * (ns + 1) / 2.0
    for i in range(1, nblk):
        # c with  c * ns - c^2 / 2 = i / nblk * area
        c = ns - (ns * ns - 2.0 * area * i / nblk) ** 0.5
        c = int(round(c / block_multiple)) * block_multiple
        if edges[-1] < c < ns:
            edges.append(c)
    edges.append(ns)
    G = L.empty((ns, ns))
    cur = torch.cuda.current_stream()
    comm = torch.cuda.Stream() if multi else None
    pieces, works = [], []
    for c0, c1 in zip(edges[:-1], edges[1:]):
        blk = dgemm_device(X_local[:, c0:c1], X_local[:, c0:], transA=True)       # (c1 - c0) x (ns - c0)
        pieces.append((c0, c1, blk))
        if multi:
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                works.append(dist.all_reduce(blk, op=dist.ReduceOp.SUM, group=group, async_op=True))
    for w in works:
        w.wait()
    if multi:
        cur.wait_stream(comm)
    for c0, c1, blk in pieces:
        G[c0:c1, c0:] = blk
        if c1 < ns:
            G[c1:, c0:c1] = blk[:, c1 - c0:].t()
    return G
