"""Generates tests/golden/* by running the UNMODIFIED reference modules.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference cannot travel to the GPU box, so its outputs do, as fixtures:

  pixel_strings.json   SHA-256 of to_dots / nominal-colour pixel strings for every
                       masked value (reference screen.py:741-789, :982-990;
                       colours.py:137-148)
  scorer_<mode>.npz    _pack, mask_and_shift_data, masked_update, apply,
                       diff_weights, compute_delta_page, byte_pair_difference
                       outputs of the reference Bitmap classes on seeded screens
                       (screen.py:207-547)
  stream_<case>.npz    opcode tuples pulled from the reference
                       Video.encode_frame (video.py:72-301) under the Movie.encode
                       schedule, with random.seed(s); np.random.seed(s), plus the
                       encoder state and the next words of both RNG streams
  helpers.npz          _make_header, _make_footer, _body, _fix_column_left/right and
                       _double_pixels outputs of the reference classes
  byte_stream.npz      Movie.emit_stream bytes for synthetic tick opcodes + the opcode
                       address table of player/iivision.dbg
  movie_<case>.npz     the byte stream of the whole Movie.encode + Movie.emit_stream loop
                       (movie.py:56-161) on synthetic frames and audio ticks, with the
                       final encoder state
  long_streams.json    SHA-256 of the opcode streams of BASELINE.json configs[3] (600-frame
                       prefix of the 6000-frame DHGR clip) and configs[4] (8 of the 64
                       clips) from the reference Video.encode_frame; --long-only, minutes
  luts.json            int(dE2000) substitution matrices from oracle/cie2000.py
                       (restated colormath; NOT reference output -- the reference
                       generator cannot run offline) together with the rows
                       recorded in SURVEY.md Appendix C

The edit-distance tables fed to the reference scorer are the oracle's
(oracle/tables.py); the scorer's arithmetic is independent of their contents.
"""

import hashlib
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from iivision_b200 import synth  # noqa: E402  (frame generator only; no CUDA)
from oracle import ref_harness, tables  # noqa: E402

STREAM_CASES = [
    # name, mode, frames, fraction, frame seed, rng seed, opcodes/frame, flip
    ("dhgr_full", "DHGR", 3, 1.0, 1, 0, 980, 292),
    ("dhgr_sparse", "DHGR", 4, 0.05, 2, 7, 980, 292),
    ("hgr_full", "HGR", 3, 1.0, 3, 0, 980, 292),
    ("hgr_sparse", "HGR", 4, 0.02, 4, 11, 600, 292),
    # one generator pulled far more than 2048 times (HGR never flips banks, so
    # --every_n_video_frames 5 at 30 fps is 2450 pulls, main.py:29, movie.py:81-95); the
    # second and third exhaust the first-pass heap and go on popping re-queued cells
    ("hgr_long_generator", "HGR", 3, 1.0, 6, 2, 2450, 292),
    ("hgr_exhaust", "HGR", 2, 1.0, 7, 3, 7000, 292),
    ("dhgr_long_generator", "DHGR", 2, 1.0, 8, 4, 5000, 10 ** 9),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def symmetric_table(mode, pid=5):
    tab, _ = tables.build_table(mode, tables.substitution_lut(pid))
    return tables.symmetrise(mode, tab)


def ref_bitmap(ns, mode, main, aux=None):
    mm = ns.screen.MemoryMap(screen_page=1, page_offset=main)
    if mode == "DHGR":
        am = ns.screen.MemoryMap(screen_page=1, page_offset=aux)
        return ns.screen.DHGRBitmap(palette=ns.palette.Palette.NTSC,
                                    main_memory=mm, aux_memory=am)
    return ns.screen.HGRBitmap(palette=ns.palette.Palette.NTSC, main_memory=mm)


def gen_pixel_strings(ns):
    out = {}
    for mode, cls, cols in (("HGR", ns.screen.HGRBitmap, ns.colours.HGRColours),
                            ("DHGR", ns.screen.DHGRBitmap, ns.colours.DHGRColours)):
        bits, n = int(cls.MASKED_BITS), int(cls.MASKED_DOTS)
        dots = np.zeros((len(cls.PHASES), 1 << bits), dtype=np.uint32)
        pix = np.zeros((len(cls.PHASES), 1 << bits, n), dtype=np.uint8)
        for o, ph in enumerate(cls.PHASES):
            for v in range(1 << bits):
                d = cls.to_dots(v, o)
                dots[o, v] = d
                pix[o, v] = ns.colours.dots_to_nominal_colour_pixel_values(
                    n, d, cols, init_phase=ph)
        out[mode] = {"dots_sha256": sha(dots), "pixels_sha256": sha(pix),
                     "dots_max": int(dots.max()), "pixels_sum": int(pix.sum()),
                     "shape": list(pix.shape)}
    with open(os.path.join(GOLDEN, "pixel_strings.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def gen_scorer(ns, mode, table):
    ref_harness.install_tables(ns, mode, {5: table})
    fr = synth.synthetic_frames(mode, 2, 1.0, seed=21)
    if mode == "HGR":
        fr[1, 0, 3, :40] = 0xFF       # palette bits set across a run
        fr[1, 0, 4, :40] = 0x80
    src = ref_bitmap(ns, mode, fr[0, 0].copy(), fr[0, 1].copy() if mode == "DHGR" else None)
    tgt = ref_bitmap(ns, mode, fr[1, 0].copy(), fr[1, 1].copy() if mode == "DHGR" else None)
    cls = type(src)
    noff = len(cls.BYTE_MASKS)
    out = {"frames": fr, "src_packed": src.packed.copy(), "tgt_packed": tgt.packed.copy()}
    rng = np.random.default_rng(5)
    width = int(cls.HEADER_BITS + cls.BODY_BITS + cls.FOOTER_BITS)
    words = rng.integers(0, 1 << width, size=(4, 128), dtype=np.uint64)
    out["words"] = words
    values = np.array([0, 1, 0x55, 0x7F, 0x80, 0xFF], dtype=np.uint8)
    out["values"] = values
    ms = np.zeros((noff,) + words.shape, np.uint64)
    mu = np.zeros((noff, len(values)) + words.shape, np.uint64)
    for o in range(noff):
        ms[o] = cls.mask_and_shift_data(words, o)
        for k, v in enumerate(values):
            mu[o, k] = cls.masked_update(o, words.copy(), np.uint8(v))
    out["mask_shift"] = ms
    out["masked_update"] = mu
    banks = (False, True) if mode == "DHGR" else (False,)
    for is_aux in banks:
        tag = "aux" if is_aux else "main"
        dw = tgt.diff_weights(src, is_aux)
        out["diff_weights_" + tag] = dw.astype(np.int32)
        cases = [(0, 0), (3, 0x55), (31, 0x7F), (17, 0x2A), (4, 0x7F if mode == "DHGR" else 0xFF),
                 (3, 0x00 if mode == "DHGR" else 0x80)]
        deltas = np.zeros((len(cases), 256), np.int32)
        for k, (page, content) in enumerate(cases):
            deltas[k] = tgt.compute_delta_page(page, np.uint8(content), dw[page, :], is_aux)
        out["delta_cases_" + tag] = np.array(cases, np.int32)
        out["delta_" + tag] = deltas
        bpd = []
        for page, off, content in ((0, 0, 5), (3, 7, 0x7F), (9, 100, 0x33), (31, 119, 1)):
            bo = cls.byte_offset(off, is_aux)
            bpd.append((bo, page, off, content, int(tgt.byte_pair_difference(
                bo, tgt.packed[page, off // 2], np.uint8(content)))))
        out["pair_difference_" + tag] = np.array(bpd, np.int64)
    # apply: a sequence of stores on the source bitmap (screen.py:256-293)
    stores = []
    rng = np.random.default_rng(9)
    for _ in range(300):
        page, off = int(rng.integers(0, 32)), int(rng.integers(0, 256))
        is_aux = bool(rng.integers(0, 2)) if mode == "DHGR" else False
        val = int(rng.integers(0, 128 if mode == "DHGR" else 256))
        stores.append((page, off, int(is_aux), val))
    stores += [(0, 0, 0, 0x7F), (0, 255, 0, 0x7F), (31, 254, 0, 1), (5, 1, 0, 0x40)]
    for page, off, is_aux, val in stores:
        src.apply(page, off, bool(is_aux), np.uint8(val))
    out["apply_stores"] = np.array(stores, np.int32)
    out["apply_packed"] = src.packed.copy()
    out["apply_main"] = src.main_memory.page_offset.copy()
    if mode == "DHGR":
        out["apply_aux"] = src.aux_memory.page_offset.copy()
    np.savez_compressed(os.path.join(GOLDEN, "scorer_%s.npz" % mode.lower()), **out)


def gen_helpers(ns):
    """Static helpers of the bitmap classes on seeded words (screen.py:650-739,
    :921-952, :295-320)."""
    out = {}
    rng = np.random.default_rng(31)
    for mode, cls in (("HGR", ns.screen.HGRBitmap), ("DHGR", ns.screen.DHGRBitmap)):
        width = int(cls.HEADER_BITS + cls.BODY_BITS + cls.FOOTER_BITS)
        words = rng.integers(0, 1 << width, size=(6, 128), dtype=np.uint64)
        other = rng.integers(0, 1 << width, size=(6, 128), dtype=np.uint64)
        bm = ref_bitmap(ns, mode, np.zeros((32, 256), np.uint8), np.zeros((32, 256), np.uint8))
        out[mode + "_words"] = words
        out[mode + "_other"] = other
        out[mode + "_header"] = cls._make_header(words.copy())
        out[mode + "_footer"] = cls._make_footer(words.copy())
        out[mode + "_fix_left"] = bm._fix_column_left(other.copy(), words.copy())
        out[mode + "_fix_right"] = bm._fix_column_right(other.copy(), words.copy())
        fr = synth.synthetic_frames(mode, 1, 1.0, seed=33)
        b2 = ref_bitmap(ns, mode, fr[0, 0].copy(), fr[0, 1].copy() if mode == "DHGR" else None)
        out[mode + "_frame"] = fr[0]
        out[mode + "_body"] = b2._body().astype(np.uint64)
    out["double_pixels"] = np.array(
        [ns.screen.HGRBitmap._double_pixels(v) for v in range(128)], np.uint64)
    np.savez_compressed(os.path.join(GOLDEN, "helpers.npz"), **out)


def gen_byte_stream(ns):
    """Movie.emit_stream (movie.py:122-161) over synthetic tick opcodes, and the player's
    opcode address table parsed from player/iivision.dbg (opcodes.py:170-217)."""
    out = {}
    tick_addr = np.zeros((32, 32), np.uint16)
    for i, tick in enumerate(range(4, 68, 2)):
        for j, page in enumerate(range(32, 64)):
            tick_addr[i, j] = ns.opcodes.TICK_OPCODES[(tick, page)]._START
    out["tick_addr"] = tick_addr
    out["ack_addr"] = np.uint32(ns.opcodes.Ack._START)
    out["terminate_addr"] = np.uint32(ns.opcodes.Terminate._START)
    rng = np.random.default_rng(41)
    for mode in ("HGR", "DHGR"):
        vm = getattr(ns.video_mode.VideoMode, mode)
        for name, n, max_out in (("short", 5, None), ("frames", 900, None),
                                 ("exact", 291 + 292, None), ("capped", 2000, 5000)):
            pages = rng.integers(32, 64, size=n)
            ticks = rng.integers(2, 34, size=n) * 2
            content = rng.integers(0, 256 if mode == "HGR" else 128, size=n)
            offs = rng.integers(0, 256, size=(n, 4))
            m = ns.movie.Movie.__new__(ns.movie.Movie)
            m.video_mode, m.max_bytes_out, m.stream_pos = vm, max_out, 0
            m.state, m.aux_memory_bank = ns.machine.Machine(), False
            ops = [ns.opcodes.Header(mode=vm)] + [
                ns.opcodes.TICK_OPCODES[(int(ticks[k]), int(pages[k]))](
                    int(content[k]), tuple(int(x) for x in offs[k])) for k in range(n)]
            data = bytes(m.emit_stream(ops))
            rec = np.zeros((n, 8), np.uint8)
            rec[:, 0], rec[:, 1], rec[:, 2:6], rec[:, 6] = pages, content, offs, 1
            key = "%s_%s" % (mode.lower(), name)
            out[key + "_records"] = rec
            out[key + "_ticks"] = ticks.astype(np.uint8)
            out[key + "_max"] = np.int64(max_out or 0)
            out[key + "_bytes"] = np.frombuffer(data, np.uint8)
    np.savez_compressed(os.path.join(GOLDEN, "byte_stream.npz"), **out)


def gen_stream(ns, case, table):
    name, mode, n_frames, fraction, fseed, seed, per_frame, flip = case
    ref_harness.install_tables(ns, mode, {5: table})
    frames = synth.synthetic_frames(mode, n_frames, fraction, seed=fseed)
    segs = synth.movie_schedule(mode, n_frames, per_frame, flip)
    vm = ns.video_mode.VideoMode.DHGR if mode == "DHGR" else ns.video_mode.VideoMode.HGR
    random.seed(seed)
    np.random.seed(seed)
    v = ns.video.Video(ns.frame_grabber.FrameGrabber(vm), ticks_per_second=14700.,
                       mode=vm, palette=ns.palette.Palette.NTSC)
    ops, real, sims = [], [], []
    for frame, is_aux, budget in segs:
        is_aux = bool(is_aux)
        tgt = ref_bitmap(ns, mode, frames[frame, 0].copy(),
                         frames[frame, 1].copy() if mode == "DHGR" else None)
        prio = v.aux_update_priority if is_aux else v.update_priority
        sims.append(float(prio.mean()))
        v.out_of_work = {True: False, False: False}
        seq = v.encode_frame(tgt, is_aux)
        for _ in range(budget):
            page, content, offs = next(seq)
            real.append(0 if v.out_of_work[is_aux] else 1)
            ops.append([int(page), int(content)] + [int(o) for o in offs])
    out = {
        "mode": mode, "n_frames": n_frames, "fraction": fraction,
        "frame_seed": fseed, "rng_seed": seed, "frames": frames,
        "segments": np.array(segs, np.int32),
        "opcodes": np.array(ops, np.uint8), "real": np.array(real, np.uint8),
        "similarity": np.array(sims, np.float64),
        "packed": v.pixelmap.packed.copy(),
        "main": v.memory_map.page_offset.copy(),
        "priority_main": v.update_priority.copy(),
        "next_python_words": np.array(
            [random.getrandbits(32) for _ in range(4)], np.uint32),
        "next_numpy_bytes": np.random.randint(0, 256, size=4).astype(np.uint32),
    }
    if mode == "DHGR":
        out["aux"] = v.aux_memory_map.page_offset.copy()
        out["priority_aux"] = v.aux_update_priority.copy()
    np.savez_compressed(os.path.join(GOLDEN, "stream_%s.npz" % name), **out)
    print("  %s: %d opcodes, %d real" % (name, len(ops), int(np.sum(real))))


MOVIE_CASES = [
    # name, mode, grabber frames, frame fraction, frame seed, rng seed, input fps, every_n,
    # audio ticks, max_bytes_out
    ("dhgr_30fps", "DHGR", 5, 1.0, 31, 3, 30, 2, 2300, None),
    ("dhgr_24fps_capped", "DHGR", 4, 0.3, 32, 4, 24, 1, 2600, 9000),
    ("hgr_30fps", "HGR", 4, 1.0, 33, 5, 30, 2, 1700, None),
    ("hgr_frames_run_out", "HGR", 2, 0.2, 34, 6, 25, 1, 2000, None),
]


def gen_movie(ns, case, table):
    """The whole of Movie.encode + Movie.emit_stream (movie.py:56-161) of the unmodified
    reference, with the two media front ends replaced by in-memory sources: an audio
    object giving sample_rate and audio_stream() (values -15..16, audio.py:84-103) and a
    frame grabber giving input_frame_rate and frames() -> (main, aux) MemoryMaps
    (frame_grabber.py:56-140).  Movie.__init__ would open the media file, so the object is
    built field by field as its __init__ does (movie.py:16-54)."""
    import contextlib
    import io
    name, mode, n_frames, fraction, fseed, seed, fps, every_n, n_ticks, max_out = case
    ref_harness.install_tables(ns, mode, {5: table})
    frames = synth.synthetic_frames(mode, n_frames, fraction, seed=fseed)
    audio = np.random.default_rng(fseed).integers(-15, 17, size=n_ticks).astype(np.int8)
    vm = getattr(ns.video_mode.VideoMode, mode)

    class Audio:
        sample_rate = 14700.

        @staticmethod
        def audio_stream():
            yield from (int(a) for a in audio)

    class Grabber:
        input_frame_rate = fps

        @staticmethod
        def frames():
            for k in range(n_frames):
                main = ns.screen.MemoryMap(screen_page=1, page_offset=frames[k, 0].copy())
                aux = (ns.screen.MemoryMap(screen_page=1, page_offset=frames[k, 1].copy())
                       if mode == "DHGR" else None)
                yield main, aux

    random.seed(seed)
    np.random.seed(seed)
    m = ns.movie.Movie.__new__(ns.movie.Movie)
    m.filename = "synthetic"
    m.every_n_video_frames = every_n
    m.max_bytes_out = max_out
    m.video_mode = vm
    m.palette = ns.palette.Palette.NTSC
    m.audio = Audio()
    m.frame_grabber = Grabber()
    m.video = ns.video.Video(m.frame_grabber, ticks_per_second=m.audio.sample_rate,
                             mode=vm, palette=m.palette)
    m.stream_pos = 0
    m.ticks = 0
    m.state = ns.machine.Machine()
    m.aux_memory_bank = False
    with contextlib.redirect_stdout(io.StringIO()):
        data = bytes(m.emit_stream(m.encode()))
    out = {
        "mode": mode, "frames": frames, "audio": audio, "rng_seed": seed,
        "input_frame_rate": np.float64(fps), "every_n_video_frames": every_n,
        "sample_rate": np.float64(14700.), "max_bytes_out": np.int64(max_out or 0),
        "bytes": np.frombuffer(data, np.uint8), "ticks_pulled": np.int64(m.ticks),
        "packed": m.video.pixelmap.packed.copy(),
        "main": m.video.memory_map.page_offset.copy(),
        "priority_main": m.video.update_priority.copy(),
    }
    if mode == "DHGR":
        out["aux"] = m.video.aux_memory_map.page_offset.copy()
        out["priority_aux"] = m.video.aux_update_priority.copy()
    np.savez_compressed(os.path.join(GOLDEN, "movie_%s.npz" % name), **out)
    print("  %s: %d bytes, %d ticks" % (name, len(data), m.ticks))


def _long_worker(job):
    """One clip through the UNMODIFIED reference Video.encode_frame under the Movie.encode
    schedule; returns cumulative SHA-256 digests of the opcode stream at checkpoints."""
    import hashlib
    kind, clip, n_frames, checkpoints = job
    ns = ref_harness.load()
    table = symmetric_table("DHGR")
    ref_harness.install_tables(ns, "DHGR", {5: table})
    if kind == "long":
        frames = synth.long_clip_frames(n_frames)
        seed = synth.LONG_CLIP["rng_seed"]
    else:
        frames = synth.batch_clip_frames(clip, n_frames)
        seed = synth.batch_clip_seeds(clip)[1]
    segs = synth.movie_schedule("DHGR", n_frames)
    vm = ns.video_mode.VideoMode.DHGR
    random.seed(seed)
    np.random.seed(seed)
    v = ns.video.Video(ns.frame_grabber.FrameGrabber(vm), ticks_per_second=14700.,
                       mode=vm, palette=ns.palette.Palette.NTSC)
    h = hashlib.sha256()
    digests = {}
    import time
    t0 = time.perf_counter()
    last_frame = 0
    for frame, is_aux, budget in segs:
        if frame != last_frame:
            if frame in checkpoints:
                digests[str(frame)] = h.hexdigest()
            last_frame = frame
        tgt = ref_bitmap(ns, "DHGR", frames[frame, 0].copy(), frames[frame, 1].copy())
        seq = v.encode_frame(tgt, bool(is_aux))
        buf = bytearray()
        for _ in range(budget):
            page, content, offs = next(seq)
            buf += bytes([int(page), int(content)] + [int(o) for o in offs])
        h.update(bytes(buf))
    digests[str(n_frames)] = h.hexdigest()
    dt = time.perf_counter() - t0
    state = hashlib.sha256()
    for a in (v.pixelmap.packed, v.memory_map.page_offset, v.aux_memory_map.page_offset,
              v.update_priority.astype(np.int32), v.aux_update_priority.astype(np.int32)):
        state.update(np.ascontiguousarray(a).tobytes())
    return {"kind": kind, "clip": clip, "n_frames": n_frames, "opcode_sha256": digests,
            "state_sha256": state.hexdigest(), "reference_frames_per_s": n_frames / dt,
            "next_python_word": int(random.getrandbits(32)),
            "next_numpy_byte": int(np.random.randint(0, 256))}


def gen_long_hashes():
    """BASELINE.json configs[3] (prefix of the 6000-frame DHGR clip) and configs[4] (8 of
    the 64 clips): SHA-256 of the opcode streams the unmodified reference emits, one
    process per clip (the reference is single-threaded, video.py:121-187)."""
    import multiprocessing as mp
    lc, bc = synth.LONG_CLIP, synth.BATCH_CLIPS
    jobs = [("long", -1, lc["golden_frames"], (1, 3, 10, 30, 100, 300))]
    jobs += [("batch", c, bc["n_frames"], (1, 4, 16)) for c in bc["golden_clips"]]
    with mp.get_context("spawn").Pool(min(len(jobs), os.cpu_count() or 1)) as pool:
        res = pool.map(_long_worker, jobs, chunksize=1)
    out = {"source": "unmodified reference Video.encode_frame (video.py:72-301) under "
                     "synth.movie_schedule (980 opcodes/frame, bank flip every 292), NTSC, "
                     "oracle tables; digest = sha256 of uint8[n][6] (page+32, content, 4 "
                     "offsets), cumulative at the frame counts given",
           "long_clip": dict(lc), "batch_clips": {k: (list(v) if isinstance(v, tuple) else v)
                                                  for k, v in bc.items()},
           "long": res[0], "batch": {str(r["clip"]): r for r in res[1:]}}
    with open(os.path.join(GOLDEN, "long_streams.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for r in res:
        print("  %s %d: %d frames at %.2f frames/s" % (
            r["kind"], r["clip"], r["n_frames"], r["reference_frames_per_s"]))


def gen_luts():
    survey_row0 = {
        "5": [0, 35, 37, 50, 38, 39, 55, 64, 31, 53, 39, 65, 66, 78, 86, 99],
        "0": [0, 44, 31, 49, 40, 24, 40, 60, 35, 57, 57, 66, 73, 101, 88, 99],
    }
    out = {"source": "oracle/cie2000.py (restated colormath 3.0.0; parity unpinned)",
           "survey_row0": survey_row0, "lut": {}, "table_sha256": {}}
    for pid in (0, 5):
        lut = tables.substitution_lut(pid)
        assert lut[0].tolist() == survey_row0[str(pid)]
        out["lut"][str(pid)] = lut.tolist()
        for mode in ("HGR", "DHGR"):
            tab, _ = tables.build_table(mode, lut)
            out["table_sha256"]["%s_%d" % (mode, pid)] = sha(tab)
    with open(os.path.join(GOLDEN, "luts.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if "--long-only" in sys.argv:
        print("long streams"); gen_long_hashes()
        return
    ns = ref_harness.load()
    if "--streams" in sys.argv:       # --streams name[,name...]: just those stream fixtures
        wanted = sys.argv[sys.argv.index("--streams") + 1].split(",")
        for mode in ("HGR", "DHGR"):
            cases = [c for c in STREAM_CASES if c[1] == mode and c[0] in wanted]
            if cases:
                table = symmetric_table(mode)
                for case in cases:
                    gen_stream(ns, case, table)
        return
    print("helpers"); gen_helpers(ns)
    print("byte stream"); gen_byte_stream(ns)
    if "--helpers-only" in sys.argv:
        return
    if "--movie-only" in sys.argv:
        for mode in ("HGR", "DHGR"):
            table = symmetric_table(mode)
            for case in MOVIE_CASES:
                if case[1] == mode:
                    gen_movie(ns, case, table)
        return
    print("pixel strings"); gen_pixel_strings(ns)
    print("luts"); gen_luts()
    for mode in ("HGR", "DHGR"):
        table = symmetric_table(mode)
        print("scorer", mode); gen_scorer(ns, mode, table)
        for case in STREAM_CASES:
            if case[1] == mode:
                gen_stream(ns, case, table)
        for case in MOVIE_CASES:
            if case[1] == mode:
                gen_movie(ns, case, table)


if __name__ == "__main__":
    main()
