"""DyYOLO — Darknet-53 whose stem and neck 1x1 convs are dynamic convolutions
(reference model/DyYOLO.py:56-164; layer type "DyConv", attention temperature from hparams)."""
from ._base import DyConvModule  # noqa: F401
from .darknet import CNNBlock, DarknetDetector, ResidualBlock, ScalePrediction  # noqa: F401


class DyYOLO(DarknetDetector):
    """`hparams` as BaselineModel plus `attn_temperature` (conf/model/dy-yolo.yaml:16)."""
    supports_dyconv = True
