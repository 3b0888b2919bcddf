"""Inference-twin classes (choijhanyangackr/yolox_infer/models/__init__.py): BN-free, raw-logit outputs."""
from .models import _InferYOLOX as YOLOX, _InferYOLOXP6 as YOLOXP6, _InferYOLOXP6v2 as YOLOXP6v2  # noqa: F401
from .models import _InferYOLOXDepthwise as YOLOXDepthwise  # noqa: F401
from .models import YOLOXHead, YOLOPAFPN, YOLOPAFPNCustomP6 as YOLOPAFPNP6  # noqa: F401
from .models import CSPDarknet, CSPDarknetCustomP6 as CSPDarknetP6  # noqa: F401
