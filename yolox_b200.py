"""Import alias: the package directory name required by the build contract contains hyphens, so
`import yolox_b200` loads `coco-dataset-based-light-weight-fast-object-detection-model_b200/`."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("coco-dataset-based-light-weight-fast-object-detection-model_b200")
sys.modules[__name__] = _pkg
