"""How the hot path is partitioned over the 1/2/4/8 GPUs of one node.

* batches of independent systems (configs 2-4): rank r owns the contiguous block
  `batch_range(nsys, r, world)`; no collective on the data path.
* one large apply / wide-symbol system (config 5): rank r owns the symbol-byte
  columns `column_slice(nbytes, r, world)` of D and Y; the coefficient matrix is
  needed by every rank and is broadcast once with NCCL (`broadcast_coefficients`).
One process per GPU, torch.distributed for the plumbing.
"""


def batch_range(nsys, rank, world):
    base, rem = divmod(nsys, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def column_slice(nbytes, rank, world, align=256):
    """[begin, end) byte columns of rank `rank`; every boundary is a multiple of
    `align` (the kernels' tile width) except the last end."""
    tiles = (nbytes + align - 1) // align
    t0, t1 = batch_range(tiles, rank, world)
    return min(t0 * align, nbytes), min(t1 * align, nbytes)


def broadcast_coefficients(C, src=0):
    """NCCL (or gloo on CPU) broadcast of the coefficient matrix to every rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(C, src=src)
    return C


def gather_ranks(rank_local, world_sizes=None):
    """All-gather of the small per-system status array (ranks) -- the only
    collective of the batch configs, and off the timed data path."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rank_local
    n = torch.tensor([rank_local.numel()], device=rank_local.device)
    sizes = [torch.zeros_like(n) for _ in range(dist.get_world_size())]
    dist.all_gather(sizes, n)
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros(mx, dtype=rank_local.dtype, device=rank_local.device)
    pad[:rank_local.numel()] = rank_local
    out = [torch.zeros_like(pad) for _ in range(dist.get_world_size())]
    dist.all_gather(out, pad)
    return torch.cat([o[:int(s)] for o, s in zip(out, sizes)])
