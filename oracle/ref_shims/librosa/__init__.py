"""Stub: the reference audio.py imports librosa at module level; unused by the oracle."""
