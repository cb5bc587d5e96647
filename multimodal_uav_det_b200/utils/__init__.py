from .datatype import BatchData, Config, DetectionResults  # noqa: F401
