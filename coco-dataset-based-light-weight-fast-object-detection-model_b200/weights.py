"""Checkpoint ingest and pruning-mask utilities (host side, torch CPU ops on weights only).

  load_checkpoint   choijhanyangackr/main.py:50-55 — flat fused state_dict (strict) or {"model": sparse COO}
  magnitude_masks   01_mask_generator.py:18-39 — global |w| threshold over non-head 4-D tensors
  merge_masks       03_jh_merge.py:43-56 intended semantics (the shipped script drops the weights)
  to_sparse_ckpt    03_jh_merge.py:66-87
The engine runs unstructured masks dense-with-zeros; see DESIGN.md for the 2:4 path status."""
from typing import Dict

import torch


def load_checkpoint(model, ckpt, sparse: bool = False):
    if isinstance(ckpt, str):
        ckpt = torch.load(ckpt, map_location="cpu")
    if sparse:
        sd = model.state_dict()
        for key, param in ckpt["model"].items():
            sd[key].copy_(param.to_dense().data)      # main.py:53-55
    else:
        model.load_state_dict(ckpt, strict=True)      # main.py:50-51
    return model


def magnitude_masks(state: Dict[str, torch.Tensor], prune_pct: float = 49.0) -> Dict[str, torch.Tensor]:
    ws = {k: v for k, v in state.items() if "head" not in k and v.ndim == 4}
    allw = torch.cat([v.detach().float().flatten() for v in ws.values()]).abs().clamp_max(1.0)
    thr = allw.sort()[0][int(len(allw) * prune_pct / 100)]
    return {k: torch.greater(v.detach().float().abs(), thr) for k, v in ws.items()}


def merge_masks(fused_state: Dict[str, torch.Tensor], masks: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = dict(fused_state)
    for k, m in masks.items():
        out[k] = fused_state[k] * m.to(fused_state[k].dtype)
    return out


def to_sparse_ckpt(fused_state: Dict[str, torch.Tensor]) -> Dict[str, Dict[str, torch.Tensor]]:
    return {"model": {k: v.to_sparse().coalesce() for k, v in fused_state.items()}}


def density(state: Dict[str, torch.Tensor]) -> float:
    ws = [v for k, v in state.items() if v.ndim == 4]
    return float(sum(int((v != 0).sum()) for v in ws)) / float(sum(v.numel() for v in ws))
