"""Stub for matplotlib (dataset/_helper.py imports pyplot at module scope)."""
