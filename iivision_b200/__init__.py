"""iivision_b200: B200-native (sm_100a) hot paths of the ii-vision transcoder.

Two paths, behind the reference's own Python API:
  * edit-distance table generation (reference transcoder/make_data_tables.py)
  * per-frame scoring / greedy delta encoding (transcoder/screen.py, video.py)
Host code is Python; every array computation runs in hand-written CUDA through
the C ABI in include/iivision_b200.h (ctypes).  There is no CPU fallback:
importing ``iivision_b200._lib`` raises if the shared library is missing.
"""

__version__ = "0.1.0"
