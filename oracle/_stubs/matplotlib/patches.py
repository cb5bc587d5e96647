def __getattr__(name):
    def _missing(*a, **k):
        raise RuntimeError("matplotlib stub")
    return _missing
