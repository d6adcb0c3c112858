"""B200-native cascade inference hot path of RapidObjectDetectionUsingCascadedCNNs.

Python host code (this package) mirrors the reference's inference API
(``InferenceCascadeApp`` / ``InferenceApp`` / ``cf`` / ``LabeledBoundingBox``) and calls
hand-written sm_100a CUDA kernels through the C ABI of ``include/rodc.h`` (``librodc.so``).
"""
from . import config as cf                                       # noqa: F401
from .app.inference_app import InferenceApp                      # noqa: F401
from .app.inference_cascade_app import InferenceCascadeApp       # noqa: F401
from .data.image_info import ArrayImageInfo, FileImageInfo       # noqa: F401
from .data.rectangles import LabeledBoundingBox, Rectangle, Window   # noqa: F401
from .network.stage_model import StageModel, xavier_cascade_models   # noqa: F401

__all__ = ["cf", "InferenceApp", "InferenceCascadeApp", "ArrayImageInfo", "FileImageInfo",
           "LabeledBoundingBox", "Rectangle", "Window", "StageModel", "xavier_cascade_models"]
