"""frame_grabber.py only touches these from FileFrameGrabber, which we never build."""


class FFmpegReader:
    def __init__(self, *a, **k):
        raise RuntimeError("skvideo stub: no ffmpeg offline")


def ffprobe(*a, **k):
    raise RuntimeError("skvideo stub: no ffmpeg offline")
