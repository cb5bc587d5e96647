"""API types of the detector hot path (reference utils/datatype.py:4-23).  Field order of
DetectionResults is (bbox, obj) — callers unpack positionally."""
from typing import List, NamedTuple, Union

import torch


class DetectionResults(NamedTuple):
    bbox: torch.Tensor
    obj: torch.Tensor


class BatchData(NamedTuple):
    image: torch.Tensor
    bbox: Union[torch.Tensor, List[torch.Tensor]]


class Config:
    """dict -> attribute view, nested (same contract as reference datatype.py:13-23); also accepts
    an existing attribute object (OmegaConf node) unchanged via `Config.wrap`."""

    def __init__(self, cfg: dict):
        for key, value in cfg.items():
            setattr(self, key, Config(value) if isinstance(value, dict) else value)

    @staticmethod
    def wrap(obj):
        return Config(obj) if isinstance(obj, dict) else obj
