"""Per-image data parallelism: every image (and every `num_samples` replica) is
independent — `utils/stable_diffusion_controlnet_inpaint.py:1540-1656` is batched-elementwise in N —
so rank r of W denoises images r, r+W, ... with replicated weights and NO collective inside the
step; one all-gather of the finished tensors at the end.  Works on NCCL (GPU) and gloo (CPU tests)."""
import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int):
    """Indices owned by `rank` (round-robin, so the per-rank counts differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


def gather_sharded(local: torch.Tensor, n_items: int, rank: int, world: int):
    """All-gather per-rank results `[n_local, ...]` (n_local = len(shard_indices(...))) into
    `[n_items, ...]` in the ORIGINAL item order on every rank.  A single collective."""
    if world == 1:
        return local
    n_max = (n_items + world - 1) // world
    pad = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    out = local.new_empty((n_items,) + tuple(local.shape[1:]))
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = bufs[r][:len(idx)]
    return out
