"""Clip sharding across the GPUs of one box.

Clips are independent in feature extraction and in CRNN inference (the only dependencies -- frame
wrap padding, the noise-floor tracker, the BiGRU -- live inside a clip), so the path shards by
contiguous clip ranges, one process per GPU, with no data-path collective.  The single exchange step
is the final gather of per-clip outputs: `gather_clip_outputs` is an
`all_gather` over the process group (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def clip_range(n_clips: int, rank: int, world_size: int):
    """Contiguous range [lo, hi) of rank `rank`: [r*C/W, (r+1)*C/W)."""
    if not (0 <= rank < world_size):
        raise ValueError('rank {} outside world of {}'.format(rank, world_size))
    return rank * n_clips // world_size, (rank + 1) * n_clips // world_size


def gather_clip_outputs(local: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """local: this rank's per-clip outputs (n_local, ...) for its clip_range -> (n_clips, ...) on every
    rank, in clip order.  Shards may differ by one clip; they are padded to the largest for the
    collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        if local.shape[0] != n_clips:
            raise ValueError('single process: expected {} clips, got {}'.format(n_clips, local.shape[0]))
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = clip_range(n_clips, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError('rank {} owns clips [{}, {}) but passed {} rows'.format(rank, lo, hi, local.shape[0]))
    n_max = max(clip_range(n_clips, r, world)[1] - clip_range(n_clips, r, world)[0] for r in range(world))
    padded = local.new_zeros((n_max,) + tuple(local.shape[1:]))
    padded[:hi - lo] = local
    out = local.new_empty((world * n_max,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    parts = []
    for r in range(world):
        rlo, rhi = clip_range(n_clips, r, world)
        parts.append(out[r * n_max:r * n_max + (rhi - rlo)])
    return torch.cat(parts, dim=0)
