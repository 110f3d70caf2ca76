"""Enum representing video encoding mode (reference transcoder/video_mode.py)."""

import enum


class VideoMode(enum.Enum):
    HGR = 0  # Hi-Res
    DHGR = 1  # Double Hi-Res
