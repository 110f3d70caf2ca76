"""Host logic: the Apple II hi-res address maps and the two memory-map classes of
iivision_b200/screen.py (reference screen.py:16-125) -- known addresses, the bijection
between visible (y, x) and non-hole (page, offset), the 512 screen holes, and, where the
reference tree is at hand, equality with its own tables."""

import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def screen():
    from iivision_b200 import screen
    return screen


def test_known_base_addresses(screen):
    want = {0: 0x2000, 1: 0x2400, 7: 0x3c00, 8: 0x2080, 63: 0x3f80, 64: 0x2028, 128: 0x2050,
            191: 0x3fd0}
    for y, addr in want.items():
        assert screen.y_to_base_addr(y) == addr
        assert screen.y_to_base_addr(y, 1) == addr + 0x2000
    assert screen.Y_TO_BASE_ADDR[0][191] == 0x3fd0 and len(screen.Y_TO_BASE_ADDR[1]) == 192


def test_visible_bytes_and_holes(screen):
    holes = screen.SCREEN_HOLES
    assert holes.shape == (32, 256) and int(holes.sum()) == 512
    # the holes are the last 8 bytes of every 128-byte half page
    assert holes.reshape(64, 128)[:, 120:].all() and not holes.reshape(64, 128)[:, :120].any()
    seen = np.zeros((32, 256), dtype=bool)
    for y in range(192):
        for x in range(40):
            p, o = int(screen.X_Y_TO_PAGE[y, x]), int(screen.X_Y_TO_OFFSET[y, x])
            assert (p + 32) * 256 + o == screen.y_to_base_addr(y) + x
            assert screen.PAGE_OFFSET_TO_Y[p, o] == y and screen.PAGE_OFFSET_TO_X[p, o] == x
            assert not seen[p, o]
            seen[p, o] = True
    assert np.array_equal(seen, ~holes)
    assert screen.ADDR_TO_COORDS[0x2000] == (0, 0, 0)
    assert screen.ADDR_TO_COORDS[0x4000 + 0x1fd0 + 39] == (1, 191, 39)
    assert len(screen.ADDR_TO_COORDS) == 2 * 192 * 40 and 0x2078 not in screen.ADDR_TO_COORDS


def test_memory_maps(screen):
    m = screen.MemoryMap(1)
    assert m.page_offset.shape == (32, 256) and m.page_offset.dtype == np.uint8
    m.write(32, 5, 0x7f)            # absolute page number
    m.write(3, 9, 0x11)             # page relative to the screen's first (reference quirk)
    assert m.page_offset[0, 5] == 0x7f and m.page_offset[3 - 32, 9] == 0x11
    flat = m.to_flat_memory_map()
    assert flat.data.shape == (8192,) and flat.data[5] == 0x7f
    flat.write(0x2000 + 300, 0x22)
    assert m.page_offset[1, 44] == 0x22                 # views of the same bytes
    back = flat.to_memory_map()
    assert back.screen_page == 1 and back.page_offset[1, 44] == 0x22
    given = np.arange(8192, dtype=np.uint8).reshape(32, 256)
    assert screen.MemoryMap(2, given).page_offset is given       # adopted, not copied
    for bad in (0, 3):
        with pytest.raises(ValueError, match="Screen page out of bounds"):
            screen.MemoryMap(bad)
        with pytest.raises(ValueError, match="Screen page out of bounds"):
            screen.FlatMemoryMap(bad)
    with pytest.raises(ValueError, match="Unexpected shape"):
        screen.MemoryMap(1, np.zeros((32, 255), dtype=np.uint8))
    with pytest.raises(ValueError, match="Unexpected shape"):
        screen.FlatMemoryMap(1, np.zeros((8191,), dtype=np.uint8))
    with pytest.raises(ValueError, match="Address out of range"):
        screen.FlatMemoryMap(1).write(0x1fff, 1)
    with pytest.raises(ValueError, match="Address out of range"):
        screen.FlatMemoryMap(2).write(0x6000, 1)


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/transcoder"),
                    reason="reference tree not present")
def test_tables_equal_the_reference(screen):
    from oracle import ref_harness
    ref = ref_harness.load().screen
    for name in ("PAGE_OFFSET_TO_X", "PAGE_OFFSET_TO_Y", "X_Y_TO_PAGE", "X_Y_TO_OFFSET",
                 "SCREEN_HOLES"):
        assert np.array_equal(getattr(screen, name), getattr(ref, name)), name
    assert screen.Y_TO_BASE_ADDR == ref.Y_TO_BASE_ADDR
    assert screen.ADDR_TO_COORDS == ref.ADDR_TO_COORDS
