"""Predict entry point: the device side of choijhanyangackr/main.py (build_yolox :31-59 and the body of
the run loop :153-203).  Host I/O around it (image folder dataset, COCO json) is the caller's, as in the
reference."""
import os
from typing import Iterable, List, Optional, Tuple

import torch

from . import dist as ydist
from . import infer, postprocess, weights


def build_yolox(cfg: dict, device="cuda"):
    """main.py:31-59: model selected by substring of cfg["model"]["type"]; sparse checkpoints densified."""
    d, w = cfg["model"]["depth"], cfg["model"]["width"]
    model_type = cfg["model"]["type"].lower()
    if "dw" in model_type:
        model = infer.YOLOXDepthwise(d, w)              # main.py:36-38
    elif "p6-v2" in model_type:
        model = infer.YOLOXP6v2(d, w, act="silu")       # main.py:39-41 (SiLU!)
    else:
        model = infer.YOLOXP6(d, w) if "p6" in model_type else infer.YOLOX(d, w)
    model.eval()
    if cfg.get("ckpt") is not None:
        weights.load_checkpoint(model, cfg["ckpt"], sparse=bool(cfg.get("sparse")))
    model = model.to(device)
    if cfg.get("half", True):
        model = model.half()
    return model


class Predictor:
    """One GPU's share of the predict loop: H2D -> (x*0.9+11.4 fused into the image read) -> forward ->
    fused decode + NMS.  No host synchronisation; returns device tensors."""

    def __init__(self, model, conf_threshold=0.001, nms_threshold=0.65, max_num_nms=5000, max_num_det=300,
                 in_scale=0.9, in_shift=11.4, use_graph=False, whole_graph=False):
        """use_graph: replay the network's launches from the engine's CUDA graph.  whole_graph: capture the WHOLE step
        (network + decode + selection + sort + NMS, ~125 launches) into one CUDA graph per input shape, so a call is one
        copy into the static input plus one graph launch -- the latency mode (bs1); the returned tensors are the graph's
        static outputs, overwritten by the next call."""
        self.model = model
        self.whole_graph = whole_graph
        self._step_graphs = {}
        self.kw = dict(conf_threshold=conf_threshold, nms_threshold=nms_threshold, max_num_nms=max_num_nms,
                       max_num_det=max_num_det)
        self.in_scale, self.in_shift, self.use_graph = in_scale, in_shift, use_graph
        self._gathers = {}
        self._peer_unavailable = False
        self.peer_check_every = 32      # fused gather: host check of the peers' status word every N sharded steps

    @torch.no_grad()
    def __call__(self, img: torch.Tensor, gather=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """img: [B,3,H,W] fp16/fp32, BGR 0-255 (host pinned or device).  -> det [B,max_det,7], count [B]."""
        dev = next(self.model.parameters()).device
        if img.device != dev:
            img = img.to(dev, non_blocking=True)                      # main.py:161
        if self.whole_graph and gather is None:
            return self._replay(img)
        return self._step(img, gather)

    def _step(self, img, gather=None):
        eng, reg8, cls = self.model.run_engine(img, self.in_scale, self.in_shift, self.use_graph)  # :164-167
        C = self.model.head.num_classes
        det, cnt, _ = postprocess.detect_main(reg8[..., :4], reg8[..., 4:5], cls[..., :C], self.model.head.hw,
                                              self.model.head.strides, gather=gather, **self.kw)       # :180-188
        return det, cnt

    def _replay(self, img):
        key = (tuple(img.shape), img.dtype)
        entry = self._step_graphs.get(key)
        if entry is None:
            static_in = img.clone()
            self._step(static_in)                       # first run: builds + tunes the engine, warms the allocator
            torch.cuda.synchronize(img.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                det, cnt = self._step(static_in)
            entry = self._step_graphs[key] = (graph, static_in, det, cnt)
        graph, static_in, det, cnt = entry
        if img.data_ptr() != static_in.data_ptr():
            static_in.copy_(img, non_blocking=True)
        graph.replay()
        return det, cnt

    def _peer_gather(self, per_rank: int, device):
        """Windows of the fused gather, created once per batch shape (collective: every rank calls it together)."""
        key = (per_rank, self.kw["max_num_det"])
        if key not in self._gathers:
            self._gathers[key] = ydist.PeerGather(per_rank, self.kw["max_num_det"], device)
        return self._gathers[key]

    def predict_sharded(self, img_global: torch.Tensor):
        """Batch-sharded multi-GPU step: this rank runs its contiguous slice, then ONE all-gather."""
        import torch.distributed as tdist
        world = tdist.get_world_size() if tdist.is_initialized() else 1
        rank = tdist.get_rank() if tdist.is_initialized() else 0
        n = img_global.shape[0]
        per = -(-n // world)
        s, e = ydist.shard_range(n, rank, world)
        mine = img_global[s:e]
        if e - s < per:                                                # pad so every rank runs the same batch
            filler = mine[-1:] if e > s else img_global[:1]            # (a rank may own no image at all when n < world)
            mine = torch.cat([mine, filler.expand(per - (e - s), -1, -1, -1)], 0)
        if world == 1:
            return self(mine)
        dev = next(self.model.parameters()).device
        fused = (tdist.get_backend() == "nccl" and self.kw["max_num_det"] > 0 and not self._peer_unavailable
                 and os.environ.get("YX_PEER_GATHER", "1") != "0")
        g = None
        if fused:
            try:
                g = self._peer_gather(per, dev)
            except RuntimeError:          # CUDA IPC not permitted here; raised on every rank together
                self._peer_unavailable = True
        if g is not None:   # the NMS kernel stores its rows into every rank's window; no collective call on the data path
            self(mine, gather=g)
            det_all, cnt_all = g.result()      # completes THIS step (views valid until the next step is issued)
            self._steps_since_check = getattr(self, "_steps_since_check", 0) + 1
            if g.step == 1 or self._steps_since_check >= self.peer_check_every:
                self._steps_since_check = 0
                g.check()                      # host sync every N steps: a late / dead peer raises instead of going unnoticed
        else:
            det, cnt = self(mine)
            det_all, cnt_all = ydist.all_gather_detections(det, cnt)
        return ydist.gathered_for_images(det_all, cnt_all, n, world, per)
