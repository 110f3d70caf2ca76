"""Synthetic screen-memory frames for tests and benchmarks.

BMP2DHR and real video are not available offline, so throughput and parity are
measured on synthetic 280x192 HGR / 560x192 DHGR screen-memory frames laid out
as the reference's MemoryMap (uint8[32][256] per bank, transcoder/screen.py:
101-125).  Each frame re-draws a fraction ``f`` of the non-hole bytes of the
previous frame with uniform values: HGR in [0, 256), DHGR in [0, 128) because
DHGR content must have the palette bit clear (transcoder/video.py:135-137).
Screen holes stay 0 (video.py:87-88 asserts this).  Uses PCG64
(``np.random.default_rng``) so the legacy MT19937 streams the encoder consumes
are not disturbed.
"""

import numpy as np


def screen_holes() -> np.ndarray:
    """bool[32][256]: cells that back no screen byte (screen.py:16-69)."""
    holes = np.ones((32, 256), dtype=bool)
    for y in range(192):
        a, d = divmod(y, 64)
        b, c = divmod(d, 8)
        base = 8192 + 1024 * c + 128 * b + 40 * a
        page, off = divmod(base, 256)
        holes[page - 32, off:off + 40] = False
    return holes


def synthetic_frames(mode: str, n_frames: int, fraction: float = 1.0,
                     seed: int = 1) -> np.ndarray:
    """Returns uint8[n_frames][banks][32][256]; banks = (main,) for HGR and
    (main, aux) for DHGR."""
    rng = np.random.default_rng(seed)
    holes = screen_holes()
    hi = 256 if mode == "HGR" else 128
    banks = 2 if mode == "DHGR" else 1
    cur = np.zeros((banks, 32, 256), dtype=np.uint8)
    out = np.zeros((n_frames, banks, 32, 256), dtype=np.uint8)
    for k in range(n_frames):
        for b in range(banks):
            sel = (rng.random((32, 256)) < fraction) & ~holes
            vals = rng.integers(0, hi, size=(32, 256), dtype=np.uint8)
            cur[b] = np.where(sel, vals, cur[b])
        out[k] = cur
    return out


def movie_schedule(mode: str, n_frames: int, opcodes_per_frame: int = 980,
                   flip_every: int = 292):
    """(frame, is_aux, budget) segments as Movie.encode drives encode_frame
    (transcoder/movie.py:67-111, 139-148): a new generator per encoded frame and,
    for DHGR, per MAIN/AUX bank flip every 2 KiB of stream = 292 tick opcodes.
    980 opcodes per frame = 14700 ticks/s at 30 fps input with
    every_n_video_frames=2 (main.py:29)."""
    segs = []
    aux = False
    count = 0
    for fr in range(n_frames):
        left = opcodes_per_frame
        while left > 0:
            if mode == "DHGR":
                room = flip_every - (count % flip_every)
                take = min(left, room)
            else:
                take = left
            segs.append((fr, int(aux), take))
            left -= take
            count += take
            if mode == "DHGR" and count % flip_every == 0:
                aux = not aux
    return segs


# ---- BASELINE.json configs[3] / configs[4]: the scorer's named workloads -------------------
# (shared by bench.py, the tests and oracle/make_golden.py so that the committed reference
# hashes describe exactly what is timed)

LONG_CLIP = {"mode": "DHGR", "n_frames": 6000, "fraction": 1.0, "frame_seed": 5,
             "rng_seed": 0, "golden_frames": 600}
BATCH_CLIPS = {"mode": "DHGR", "n_clips": 64, "n_frames": 32, "fraction": 1.0,
               "golden_clips": (0, 9, 18, 27, 36, 45, 54, 63)}


def long_clip_frames(n_frames: int = None) -> np.ndarray:
    """configs[3]: the 6000-frame 560x192 main+aux DHGR clip (every frame re-draws every
    non-hole byte).  The generator is sequential, so a shorter call returns a prefix."""
    c = LONG_CLIP
    return synthetic_frames(c["mode"], n_frames or c["n_frames"], c["fraction"],
                            seed=c["frame_seed"])


def batch_clip_seeds(clip: int):
    """configs[4]: (frame seed, RNG seed) of clip number ``clip`` of the 64."""
    return 1000 + clip, 100 + clip


def batch_clip_frames(clip: int, n_frames: int = None) -> np.ndarray:
    c = BATCH_CLIPS
    return synthetic_frames(c["mode"], n_frames or c["n_frames"], c["fraction"],
                            seed=batch_clip_seeds(clip)[0])


def opcode_digest(opcodes6: np.ndarray) -> str:
    """SHA-256 of an opcode stream as uint8[n][6] = (page + 32, content, 4 offsets)."""
    import hashlib
    a = np.ascontiguousarray(np.asarray(opcodes6)[:, :6], dtype=np.uint8)
    return hashlib.sha256(a.tobytes()).hexdigest()
