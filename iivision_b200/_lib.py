"""ctypes binding of libiivision_b200.so (the C ABI in include/iivision_b200.h).

There is no CPU fallback: if the shared library is missing this module raises
ImportError telling the user to build it (``python -m iivision_b200._build``).
torch is used by the callers for device buffers and streams only; nothing here
takes a torch type.
"""

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IIV_LIB_PATH") or os.path.join(_HERE, "libiivision_b200.so")

MODE_HGR, MODE_DHGR = 0, 1
LAYOUT_TRIANGULAR, LAYOUT_SYMMETRIC = 0, 1
ALGO_AUTO, ALGO_CHAIN, ALGO_TREE, ALGO_SPLIT = 0, 1, 2, 3
CLIP_STATE_FIELDS = 8

c_int, c_size_t, c_void_p = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p
c_u32, c_u8 = ctypes.c_uint32, ctypes.c_uint8

# name -> (restype, argtypes); mirrors include/iivision_b200.h one to one.
PROTOTYPES = {
    "iiv_last_error": (ctypes.c_char_p, []),
    "iiv_version": (c_int, []),
    "iiv_set_l2_fetch_granularity": (c_int, [c_size_t]),
    "iiv_get_l2_fetch_granularity": (c_size_t, []),
    "iiv_fill_probe": (c_int, [c_void_p, c_size_t, c_int, c_void_p]),
    "iiv_mode_info": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "iiv_lut_cie2000": (c_int, [c_void_p, c_void_p]),
    "iiv_lut_cie2000_f64": (c_int, [c_void_p, c_void_p]),
    "iiv_all_dots": (c_int, [c_int, c_void_p, c_void_p]),
    "iiv_all_pixel_strings": (c_int, [c_int, c_void_p, c_void_p]),
    "iiv_table_generate": (c_int, [c_int, c_void_p, c_void_p, c_u32, c_u32,
                                   c_int, c_int, c_void_p]),
    "iiv_table_split_windows": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "iiv_table_generate_scatter": (c_int, [c_int, c_void_p, c_void_p, c_int,
                                           c_int, c_void_p, c_u32, c_u32, c_int,
                                           c_void_p]),
    "iiv_table_download": (c_int, [c_int, c_void_p, c_void_p, c_u32, c_u32, c_int, c_int,
                                   c_void_p]),
    "iiv_deflate_block_bytes": (c_size_t, []),
    "iiv_deflate_block_stride": (c_size_t, []),
    "iiv_deflate_survey": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                   c_void_p]),
    "iiv_deflate_encode": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "iiv_deflate_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "iiv_table_symmetrise": (c_int, [c_int, c_void_p, c_void_p]),
    "iiv_pack": (c_int, [c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_int,
                         c_void_p]),
    "iiv_mask_and_shift": (c_int, [c_int, c_int, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "iiv_masked_update": (c_int, [c_int, c_int, c_void_p, c_u8, c_void_p,
                                  c_size_t, c_void_p]),
    "iiv_column_part": (c_int, [c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "iiv_fix_column": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    "iiv_fix_array_neighbours": (c_int, [c_int, c_int, c_void_p, c_int,
                                         c_void_p]),
    "iiv_diff_weights": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int,
                                 c_void_p, c_void_p, c_int, c_void_p]),
    "iiv_score_frames": (c_int, [c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_size_t,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_void_p]),
    "iiv_score_factors_bytes": (c_size_t, [c_int]),
    "iiv_score_factors": (c_int, [c_int, c_void_p, c_void_p, c_void_p]),
    "iiv_score_frames_factored": (c_int, [c_int, c_void_p, c_size_t, c_void_p, c_void_p,
                                          c_size_t, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_void_p]),
    "iiv_score_factor_segments": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "iiv_diff_weights_page": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p, c_int, c_void_p]),
    "iiv_compute_delta_page": (c_int, [c_int, c_int, c_void_p, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "iiv_delta_rows": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_int, c_void_p]),
    "iiv_byte_pair_difference": (c_int, [c_int, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_size_t,
                                         c_void_p]),
    "iiv_apply": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                          c_void_p]),
    "iiv_clip_state_bytes": (c_size_t, []),
    "iiv_clip_state_layout": (c_int, [c_void_p]),
    "iiv_encode_clips": (c_int, [c_int, c_int, c_void_p, c_size_t, c_void_p,
                                 c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "iiv_encode_clips_planned": (c_int, [c_int, c_int, c_void_p, c_size_t, c_void_p,
                                         c_void_p, c_int, c_void_p, c_void_p, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p]),
    "iiv_encode_generator": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "iiv_event_create": (c_void_p, []),
    "iiv_event_wait": (c_int, [c_void_p]),
    "iiv_event_destroy": (c_int, [c_void_p]),
    "iiv_mt_draw": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "iiv_stream_length": (c_size_t, [c_size_t, c_int]),
    "iiv_stream_ticks_within": (c_size_t, [c_size_t, c_size_t]),
    "iiv_emit_stream": (c_int, [c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_u32,
                                c_u32, c_void_p, c_size_t, c_void_p, c_void_p]),
    "iiv_string_distance": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_void_p, c_void_p]),
}


class IIVError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libiivision_b200 error %d: %s" % (code, message))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found: build it with `python -m iivision_b200._build` "
            "(nvcc, sm_100a). iivision_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)     # AttributeError if the ABI is incomplete
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise IIVError(rc, lib.iiv_last_error().decode("utf-8", "replace"))


def mode_info(mode: int):
    bits, dots, noff = c_int(), c_int(), c_int()
    phases = (c_int * 4)()
    check(lib.iiv_mode_info(mode, ctypes.byref(bits), ctypes.byref(dots),
                            ctypes.byref(noff), phases))
    return bits.value, dots.value, noff.value, list(phases)[:noff.value]


def clip_state_layout():
    offs = (c_size_t * CLIP_STATE_FIELDS)()
    check(lib.iiv_clip_state_layout(offs))
    return int(lib.iiv_clip_state_bytes()), [int(x) for x in offs]
