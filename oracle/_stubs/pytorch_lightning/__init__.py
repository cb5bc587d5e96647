"""Minimal stand-in for pytorch_lightning so the reference model files import in a
container without Lightning.  Test infrastructure only."""
import torch


class LightningModule(torch.nn.Module):
    def log(self, *a, **k):
        return None

    def log_dict(self, *a, **k):
        return None


class Trainer:  # pragma: no cover - never driven here
    def __init__(self, *a, **k):
        raise RuntimeError("pytorch_lightning stub: Trainer is not available")


def seed_everything(seed, workers=False):
    torch.manual_seed(seed)
    return seed
