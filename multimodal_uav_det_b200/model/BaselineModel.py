"""BaselineModel — YOLOv3 / Darknet-53 detector (reference model/BaselineModel.py:56-144),
constructed exactly like `train.py:24-25` does: `BaselineModel(hparams=hparams)`."""
from .darknet import CNNBlock, DarknetDetector, ResidualBlock, ScalePrediction  # noqa: F401


class BaselineModel(DarknetDetector):
    """`hparams`: anchors, head_scales, lr, lr_scheduler, loss_balancing, bbox_loss_fn, optim,
    layer_config (conf/model/baseline.yaml).  forward(x) -> List[DetectionResults]."""
    supports_dyconv = False
