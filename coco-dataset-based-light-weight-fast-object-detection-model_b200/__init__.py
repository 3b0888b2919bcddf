"""yolox_b200 — B200-native (sm_100a) inference engine for the reference's YOLOX / YOLOX-P6 hot path.

Public surface mirrors the reference (see DESIGN.md / INTEGRATION.md):
  infer.{YOLOX, YOLOXP6}                      == choijhanyangackr/yolox_infer/models
  models.{YOLOX, YOLOXCustomP6, YOLOPAFPN, YOLOPAFPNCustomP6, YOLOXHead, YOLOXHeadCustom, fuse_model}
                                              == yolox/models, yolox/utils/model_utils.fuse_model
  postprocess.{yolox_generate_grid, yolox_postprocess_output_torch_batch, yolox_nms_torch_batch,
               postprocess, decode_outputs, detect_main}
Importing the package loads lib/libyolox_b200.so (building it with nvcc if missing); there is no
fallback compute path.
"""
from . import _capi

_capi.load()  # fail loudly if the native library cannot be built / loaded

from . import blocks, dist, models, plan, postprocess, weights  # noqa: E402
from . import infer, predict  # noqa: E402
from . import cocoeval, evaluator  # noqa: E402  (the evaluation loop around the hot path)
from . import io  # noqa: E402  (rows next to the hot path: device-side pre-processing, COCO records)
from .models import YOLOX, YOLOXCustomP6, YOLOPAFPN, YOLOPAFPNCustomP6, YOLOXHead, YOLOXHeadCustom, fuse_model  # noqa: E402,F401
from .postprocess import (decode_outputs, detect_main, postprocess as postprocess_fn, yolox_generate_grid,  # noqa: E402,F401
                          yolox_nms_torch_batch, yolox_postprocess_output_torch_batch)

__all__ = ["infer", "models", "postprocess", "plan", "blocks", "io"]
