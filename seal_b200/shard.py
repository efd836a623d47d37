"""Batch sharding across GPUs: every ciphertext of a batch is independent, so rank r of W owns one
contiguous slice and there is no collective in the data path; the only exchange is an end-of-run gather of one 64-bit
digest per ciphertext (full outputs of the headline config, 266 GB, would not fit one GPU)."""


def shard_range(batch, rank, world):
    """contiguous [lo, hi) slice of a global batch owned by `rank` (sizes differ by at most one)"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def digest(out_slab):
    """one wrapping 64-bit sum per ciphertext of a [B, ...] int64 slab (torch tensor, any device)"""
    return out_slab.reshape(out_slab.shape[0], -1).sum(dim=1)


def gather_digests(local_digest, rank, world, group=None):
    """gather the per-ciphertext digests of all ranks on rank 0 (NCCL on GPUs, gloo in the CPU tests).
    Shards may differ in length by one, so lengths are exchanged first.  Returns the concatenated global digest on
    rank 0 and None elsewhere."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local_digest
    n_local = torch.tensor([local_digest.numel()], dtype=torch.int64, device=local_digest.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    m = int(max(int(s.item()) for s in sizes))
    padded = torch.zeros(m, dtype=local_digest.dtype, device=local_digest.device)
    padded[: local_digest.numel()] = local_digest
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)])
