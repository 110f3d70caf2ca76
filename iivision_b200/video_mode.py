"""The two Apple II graphics modes the transcoder targets.  API-compatible with the
reference's ``video_mode.VideoMode`` (members and values): the value doubles as the mode
byte of the stream header and as IIV_MODE_* of the C ABI."""

import enum


class VideoMode(enum.Enum):
    HGR = 0
    DHGR = 1

    @property
    def banks(self) -> int:
        """Memory banks a frame occupies: main only, or main + aux."""
        return 2 if self is VideoMode.DHGR else 1
