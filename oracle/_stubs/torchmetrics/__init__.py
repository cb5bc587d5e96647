"""Stub: reference imports torchmetrics.detection.MeanAveragePrecision at module scope
(utils/metrics.py:4) but only uses it when return_ap=True (never in a live path)."""
