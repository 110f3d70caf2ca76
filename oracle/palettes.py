"""sRGB triples of the two palettes, indexed by nominal colour value 0..15.

TEST INFRASTRUCTURE ONLY.  Restates reference transcoder/palette.py:37-54
(NTSC, id 5) and :61-78 (IIGS, id 0); index = HGRColours.value
(colours.py:27-42).  GREY1 (5) and GREY2 (10) are identical in NTSC.
"""

import numpy as np

PALETTE_IDS = {"IIGS": 0, "NTSC": 5}

RGB = {
    5: np.array([
        (0, 0, 0), (148, 12, 125), (32, 54, 212), (188, 55, 255),
        (51, 111, 0), (126, 126, 126), (7, 168, 225), (158, 172, 255),
        (99, 77, 0), (249, 86, 29), (126, 126, 126), (255, 129, 236),
        (67, 200, 0), (221, 206, 23), (93, 248, 133), (255, 255, 255),
    ], dtype=np.uint8),
    0: np.array([
        (0, 0, 0), (221, 0, 51), (0, 0, 153), (221, 0, 221),
        (0, 119, 0), (85, 85, 85), (34, 34, 255), (102, 170, 255),
        (136, 85, 34), (255, 102, 0), (170, 170, 170), (255, 153, 136),
        (0, 221, 0), (255, 255, 0), (0, 255, 153), (255, 255, 255),
    ], dtype=np.uint8),
}
