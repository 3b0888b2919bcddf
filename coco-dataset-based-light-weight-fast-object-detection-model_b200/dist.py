"""Multi-GPU plumbing for the hot path: images are sharded by batch across ranks (one process per GPU,
weights replicated, NO collective on the forward pass) and the fixed-shape detections are collected with
ONE all-gather.  Replaces the reference's pickled Gloo object gather
(yolox/evaluators/coco_evaluator.py:127, yolox/utils/dist.py:224-265)."""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of the global batch owned by `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_images, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_detections(det: torch.Tensor, cnt: torch.Tensor) -> torch.Tensor:
    """det [B,R,7] fp32 + cnt [B] int32 -> one [B, R*7+1] fp32 wire tensor (count in the last column;
    exact for counts < 2^24)."""
    B = det.shape[0]
    return torch.cat([det.reshape(B, -1), cnt.reshape(B, 1).to(det.dtype)], dim=1).contiguous()


def unpack_detections(packed: torch.Tensor, rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
    n = packed.shape[0]
    return packed[:, :rows * 7].reshape(n, rows, 7), packed[:, rows * 7].round().to(torch.int32)


def all_gather_detections(det: torch.Tensor, cnt: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One collective: every rank receives the detections of the whole global batch, rank-major
    ([world*B, R, 7], [world*B]).  All ranks must pass the same B (pad the last shard)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return det, cnt
    world = dist.get_world_size(group)
    packed = pack_detections(det, cnt)
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    return unpack_detections(torch.cat(out, 0), det.shape[1])


def gathered_for_images(det_all, cnt_all, n_images: int, world: int, per_rank: int):
    """Drop the padding rows added so every rank ran the same batch size."""
    keep = []
    for r in range(world):
        s, e = shard_range(n_images, r, world)
        keep.extend(range(r * per_rank, r * per_rank + (e - s)))
    idx = torch.tensor(keep, device=det_all.device, dtype=torch.long)
    return det_all[idx], cnt_all[idx]
