"""Apple II nominal display colours (reference transcoder/colours.py).

Host-side scalar helpers with the reference's names and signatures.  The bulk
evaluation over every masked value happens on the device
(``ops.all_pixel_strings``); these exist so code written against the reference
module keeps working, and they are checked against the device kernels in
tests/.  A colour is the 4-bit dot window rotated left by the NTSC phase
(colours.py:100-134); both enums enumerate 0..15 and act as validators only.
"""

import enum
import functools
from typing import Tuple, Type


class NominalColours(enum.Enum):
    pass


def ror(int4: int, howmany: int) -> int:
    """Rotate-right an int4 some number of times."""
    k = howmany % 4
    return ((int4 >> k) | (int4 << (4 - k))) & 0xF


def rol(int4: int, howmany: int) -> int:
    """Rotate-left an int4 some number of times."""
    k = howmany % 4
    return ((int4 << k) | (int4 >> (4 - k))) & 0xF


# Colour names in order of their 4-bit HGR dot pattern 0b0000..0b1111
# (colours.py:27-42).  The DHGR patterns are the same sixteen names with every
# pattern rotated right by one dot: the colour reference signal is one tick out
# of phase between the two modes (colours.py:56-71).
_NAMES_BY_HGR_PATTERN = (
    "BLACK", "MAGENTA", "DARK_BLUE", "VIOLET", "DARK_GREEN", "GREY1",
    "MED_BLUE", "LIGHT_BLUE", "BROWN", "ORANGE", "GREY2", "PINK", "GREEN",
    "YELLOW", "AQUA", "WHITE",
)

HGRColours = NominalColours(
    "HGRColours", [(n, v) for v, n in enumerate(_NAMES_BY_HGR_PATTERN)],
    module=__name__)
DHGRColours = NominalColours(
    "DHGRColours",
    [(n, ror(v, 1)) for v, n in enumerate(_NAMES_BY_HGR_PATTERN)],
    module=__name__)


@functools.lru_cache(None)
def dots_to_nominal_colour_pixel_values(
        num_bits: int, dots: int, colours: Type[NominalColours] = HGRColours,
        init_phase: int = 1) -> Tuple[int, ...]:
    """Sequence of num_bits nominal colour values via sliding 4-bit window."""
    return tuple(rol((dots >> t) & 0xF, (init_phase + t) & 3)
                 for t in range(num_bits))


@functools.lru_cache(None)
def dots_to_nominal_colour_pixels(
        num_bits: int, dots: int, colours: Type[NominalColours],
        init_phase: int = 1) -> Tuple[NominalColours, ...]:
    """Sequence of num_bits nominal colour pixels via sliding 4-bit window."""
    return tuple(colours(v) for v in dots_to_nominal_colour_pixel_values(
        num_bits, dots, colours, init_phase))
