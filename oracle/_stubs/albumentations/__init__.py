"""Stub: dataset/_helper.py imports albumentations at module scope; the model files drag
the dataset package in (BaselineModel.py:6).  Nothing here is ever called."""


def __getattr__(name):
    def _missing(*a, **k):
        raise RuntimeError(f"albumentations stub: {name} is not available")
    return _missing
