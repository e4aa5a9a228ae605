"""Multi-GPU: videos are independent units -- each rank runs the full sampler on its slice,
no collective on the data path; one all-gather of the predicted frames at the end."""
import torch
import torch.distributed as dist


def shard_range(n_videos, rank, world):
    """Contiguous, balanced slice [lo, hi) of the video indices for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_videos, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_predictions(local, n_videos):
    """All-gather per-rank predictions (b_r, 3, T, H, W) into (n_videos, 3, T, H, W) in rank order.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_videos, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty(world * mx, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)
