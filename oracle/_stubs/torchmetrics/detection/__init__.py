class MeanAveragePrecision:
    def __init__(self, *a, **k):
        raise RuntimeError("torchmetrics stub: mAP is out of scope (SURVEY.md §2)")
