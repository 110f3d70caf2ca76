"""Stub: the reference audio.py imports audioread at module level; unused by the oracle."""
