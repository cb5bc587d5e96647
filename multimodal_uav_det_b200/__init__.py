"""multimodal_uav_det_b200 — B200-native detector hot path of alfialdo/multimodal-uav-det.

Public surface mirrors the reference: `model.BaselineModel`, `model.DyYOLO`,
`model.DySOEM_SimFPN`, `model.RTMUAVDet`, `utils.datatype`, `utils.postprocess`, plus `ops`
(tensor-level access to the C-ABI kernels) and `inference.detect` (decode + NMS)."""
__version__ = "0.1.0"
