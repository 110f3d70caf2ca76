"""Import stub for scikit-video (reference transcoder/frame_grabber.py:10)."""
