"""RGB palette values for rendering NominalColour pixels
(reference transcoder/palette.py).

colormath is not a dependency here: the CIE2000 pipeline runs on the device
(``ops.lut_cie2000``), so an RGB entry is a plain (r, g, b) triple of 0..255
ints rather than a colormath sRGBColor.
"""

import enum
from typing import Dict, Tuple, Type

import numpy as np

from .colours import HGRColours

RGB = Tuple[int, int, int]


def rgb(r, g, b) -> RGB:
    return (int(r), int(g), int(b))


class Palette(enum.Enum):
    """BMP2DHR palette numbers."""

    UNKNOWN = -1
    IIGS = 0
    NTSC = 5


class BasePalette:
    ID = Palette.UNKNOWN  # type: Palette

    # Palette RGB map
    RGB = {}  # type: Dict[HGRColours, RGB]

    @classmethod
    def rgb_by_value(cls) -> np.ndarray:
        """uint8[16][3] indexed by nominal colour value (HGRColours.value)."""
        out = np.zeros((16, 3), dtype=np.uint8)
        for colour, triple in cls.RGB.items():
            out[colour.value] = triple
        return out


# sRGB triples indexed by nominal colour value 0..15 (= HGRColours.value):
# BLACK MAGENTA DARK_BLUE VIOLET DARK_GREEN GREY1 MED_BLUE LIGHT_BLUE
# BROWN ORANGE GREY2 PINK GREEN YELLOW AQUA WHITE.
# NTSC = BMP2DHGR's default NTSC palette (palette.py:37-54); note GREY1 == GREY2.
# IIGS = BMP2DHGR's KEGS32 palette (palette.py:61-78).
_NTSC_BY_VALUE = (
    (0, 0, 0), (148, 12, 125), (32, 54, 212), (188, 55, 255),
    (51, 111, 0), (126, 126, 126), (7, 168, 225), (158, 172, 255),
    (99, 77, 0), (249, 86, 29), (126, 126, 126), (255, 129, 236),
    (67, 200, 0), (221, 206, 23), (93, 248, 133), (255, 255, 255),
)
_IIGS_BY_VALUE = (
    (0, 0, 0), (221, 0, 51), (0, 0, 153), (221, 0, 221),
    (0, 119, 0), (85, 85, 85), (34, 34, 255), (102, 170, 255),
    (136, 85, 34), (255, 102, 0), (170, 170, 170), (255, 153, 136),
    (0, 221, 0), (255, 255, 0), (0, 255, 153), (255, 255, 255),
)


def _by_colour(triples) -> Dict[HGRColours, RGB]:
    return {HGRColours(v): rgb(*t) for v, t in enumerate(triples)}


class NTSCPalette(BasePalette):
    ID = Palette.NTSC
    RGB = _by_colour(_NTSC_BY_VALUE)


class IIGSPalette(BasePalette):
    ID = Palette.IIGS
    RGB = _by_colour(_IIGS_BY_VALUE)


PALETTES = {
    Palette.IIGS: IIGSPalette,
    Palette.NTSC: NTSCPalette,
}  # type: Dict[Palette, Type[BasePalette]]
