#!/usr/bin/env python
"""bench.py -- headline benchmark of the ii-vision hot paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): HGR NTSC edit-distance table generation,
2 byte offsets x 2^28 entries of uint16 (1 GiB), in the symmetric in-memory
layout Bitmap.edit_distances hands to the scorer (reference screen.py:343-367).
A step = one full table.  "entries" = elements of the uint16[n_off][4**bits]
array delivered (the same count for every arm and layout).

  value      table entries/s, table left resident in HBM (device-timed)
  e2e        same through the reference-facing call
             make_data_tables.compute_edit_distance(edp, HGRBitmap, HGRColours):
             host parameters in, host numpy array out (D2H inside the timed region)
  roofline   the generator kernel against HBM write bandwidth: 2 B/entry
  cpu_baseline / --impl reference
             the reference's CPU algorithm (oracle/tables_oracle.c faithful port:
             per pair the full (n+2)^2 float64 Damerau-Levenshtein DP of
             weighted_levenshtein.dam_lev, j < i only, make_data_tables.py:143-172)
             on all host cores over a bounded sample of row blocks
  scorer     secondary figures for the second hot path (DHGR frames/s), N=1 only

N > 1 (one process per GPU):
  value      the table resident on EVERY GPU, as the per-GPU scorers need it: each rank
             generates its own replica, no collective (one GPU makes the table faster
             than NVLink can deliver (N-1)/N of it); entries counted per replica
  e2e        ONE host array: rows shard over the ranks, every rank copies its block into
             its page-locked, NUMA-local slice of a shared host array over its own PCIe
             link (parallel.compute_edit_distance_sharded)
  alt_exchange   sharded generation + NVLink exchange (peer stores / multicast / NCCL
             all-gather), each checked against the local replica
  scorer     BASELINE.json configs[4]: 64 distinct clips sharded clip-per-GPU
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODE = "HGR"
PALETTE_ID = 5  # NTSC
BITS, N_OFF, N_DOTS = 14, 2, 18
ENTRIES = N_OFF * (1 << (2 * BITS))
METRIC = "edit-distance table entries/s (HGR NTSC)"
UNIT = "entries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-scorer", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scorer-only", action="store_true",
                    help="profiling aid: run just the second hot path's figures")
    return ap.parse_args()


# ---- CPU arm: the reference's algorithm on the host cores --------------------------

def cpu_run_sample(tables, lut, out, n_rows):
    """n_rows rows strided evenly over the whole index range, so the sample's
    fill (j < i) is that of the whole triangle.  Returns (array entries
    delivered, dam_lev evaluations, seconds)."""
    n = 1 << BITS
    step = max(1, n // n_rows)
    t0 = time.perf_counter()
    _, evals = tables.build_table(MODE, lut, step // 2, n, faithful=True,
                                  triangular=True, out=out, row_step=step)
    dt = time.perf_counter() - t0
    rows = len(range(step // 2, n, step))
    return rows * n * N_OFF, evals, dt


def _cpu_rows_for(tables, lut, out, target_seconds, cores):
    probe = max(8, cores)
    _, _, dt = cpu_run_sample(tables, lut, out, probe)
    rows = int(probe * target_seconds / max(dt, 1e-3))
    return max(probe, min(1 << BITS, rows // cores * cores))


def cpu_baseline(target_seconds=12.0):
    import numpy as np
    from oracle import tables
    lut = tables.substitution_lut(PALETTE_ID)
    out = np.zeros((N_OFF, 1 << (2 * BITS)), dtype=np.uint16)
    cores = tables.max_threads()
    rows = _cpu_rows_for(tables, lut, out, target_seconds, cores)
    entries, evals, dt = cpu_run_sample(tables, lut, out, rows)
    return {
        "value": entries / dt, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "%d rows strided evenly over the table (%d dam_lev evaluations, "
                  "full (n+2)^2 float64 DP each, j<i only) in %.1f s; the whole HGR "
                  "table extrapolates to %.0f s on these cores" % (
                      rows, evals, dt, ENTRIES / (entries / dt)),
        "dam_lev_per_s": evals / dt,
    }


def run_reference(args, rank):
    if rank != 0:
        return
    import numpy as np
    from oracle import tables
    lut = tables.substitution_lut(PALETTE_ID)
    out = np.zeros((N_OFF, 1 << (2 * BITS)), dtype=np.uint16)
    cores = tables.max_threads()
    rows = _cpu_rows_for(tables, lut, out, 2.0, cores)   # ~2 s of CPU work per step
    for _ in range(args.warmup):
        cpu_run_sample(tables, lut, out, rows)
    t0 = time.perf_counter()
    entries = evals = 0
    for _ in range(args.steps):
        e, v, _ = cpu_run_sample(tables, lut, out, rows)
        entries += e
        evals += v
    dt = time.perf_counter() - t0
    value = entries / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (palette constants; no external data)",
        "config": workload_config(args.gpus),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "each step: %d rows strided evenly over the table, full "
                      "(n+2)^2 float64 dam_lev per pair (j<i), all host threads"
                      % rows},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {
        "workload": "HGR NTSC edit-distance table (make_data_tables.compute_edit_distance): "
                    "2 offsets x 2^28 uint16 entries = 1 GiB, 18-pixel strings",
        "layout": "symmetric (Bitmap.edit_distances form) for value; reference "
                  "lower-triangular array for e2e",
        "entries_per_step": ENTRIES * n_gpus,
        "l2": "each step writes 1 GiB per GPU (> 126 MB L2); no explicit flush",
        "parallelism": (
            "one GPU" if n_gpus == 1 else
            "value: the table resident on every GPU (what the per-GPU scorers gather from) = "
            "%d replicas, each rank generates its own, no collective; entries counted per "
            "replica (weak scaling).  e2e: ONE host array, rows sharded over the %d ranks, "
            "each rank copies its block home over its own PCIe link.  alt_exchange: sharded "
            "generation + NVLink exchange variants, slower than replicas" % (n_gpus, n_gpus)),
    }


# ---- clocks ------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.15:
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, parts[4:8]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- our arm ----------------------------------------------------------------------------------

def ncu_traffic():
    """DRAM bytes per launch of the generator from the committed ncu capture (N = 1,
    whole HGR table), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "split_kernel_traffic.json")) as f:
            return float(json.load(f)["traffic_bytes_per_launch"])
    except (OSError, ValueError, KeyError):
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, ValueError):
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


def golden_streams():
    """Reference digests of the named scorer workloads (tests/golden/long_streams.json,
    written by oracle/make_golden.py --long-only from the unmodified reference)."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "long_streams.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def _encode_setup(torch, ops, clips, seeds_unused=None):
    """clips uint8[n_clips, n_frames, 2, 32, 256] -> (target memory, packed targets)."""
    import numpy as np
    tmem = torch.from_numpy(np.ascontiguousarray(clips)).cuda()
    n_clips, n_frames = tmem.shape[0], tmem.shape[1]
    flat = tmem.view(-1, 2, 32, 256)
    tpacked = ops.pack("DHGR", flat[:, 0].contiguous(), flat[:, 1].contiguous()).view(
        n_clips, n_frames, 32, 128)
    return tmem, tpacked


def _timed_encode(torch, ops, tmem, tpacked, plan, table, seeds, reps, warm=2):
    """Median CUDA-event time of `reps` encode launches from fresh states; returns
    (ms, all ms sorted, opcodes of the last run, seg_info, states)."""
    n_clips = tmem.shape[0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    # outputs are allocated once: a fresh cudaMalloc between the two events would be
    # charged to the kernel
    opc = torch.empty((n_clips, plan.total, 8), dtype=torch.uint8, device="cuda")
    info = torch.zeros((n_clips, len(plan), 8), dtype=torch.int64, device="cuda")
    times = []
    st = None
    for r in range(reps + warm):
        st = ops.seed_clip_states(ops.new_clip_states(n_clips), seeds)
        torch.cuda.synchronize()
        ev[0].record()
        ops.encode_clips("DHGR", st, tmem, tpacked, plan, table, opcodes=opc, seg_info=info)
        ev[1].record()
        torch.cuda.synchronize()
        if r >= warm:
            times.append(ev[0].elapsed_time(ev[1]))
    times.sort()
    return times[len(times) // 2], times, opc, info, st


def _trace(info):
    inf = info.cpu().numpy()
    cyc = inf[0].sum(axis=0)
    return {"slowest_clip_sm_cycles": int((inf[:, :, 4] + inf[:, :, 5]).sum(axis=1).max()),
            "opcodes": int(cyc[0]), "cycles_score_heapify": int(cyc[4]),
            "cycles_opcode_loop": int(cyc[5]), "cycles_wait_rows": int(cyc[6]),
            "cycles_wait_mt_applier": int(cyc[7])}


def phase_a_figures(torch, ops, table, lut):
    """Scoring prologue of Video._index_changes (video.py:109-116) batched over DISTINCT
    frames: one launch packs every target, scores both banks against the previous frame's
    bitmap and folds the priorities.  Two ways to the same bits: iiv_score_frames_factored
    (edit-distance entries evaluated from factor tables in shared memory -- the figure
    quoted) and iiv_score_frames (entries gathered from the 512 MiB table in HBM)."""
    from iivision_b200 import synth
    nb = int(os.environ.get("IIV_BENCH_SCORE_FRAMES", "1024"))
    factors = ops.score_factors("DHGR", lut)

    def timed(fraction, **dist):
        fr = synth.synthetic_frames("DHGR", nb + 1, fraction, seed=1)
        d = torch.from_numpy(fr).cuda()
        src = ops.pack("DHGR", d[:nb, 0].contiguous(), d[:nb, 1].contiguous())
        tgt = d[1:].contiguous()
        prio = torch.zeros((nb, 2, 32, 256), dtype=torch.int32, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            ops.score_frames("DHGR", src, tgt, priority=prio, **dist)
        torch.cuda.synchronize()
        reps = 20
        ev[0].record()
        for _ in range(reps):
            ops.score_frames("DHGR", src, tgt, priority=prio, **dist)
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / reps

    def same_bits():
        fr = synth.synthetic_frames("DHGR", 65, 1.0, seed=2)
        d = torch.from_numpy(fr).cuda()
        src = ops.pack("DHGR", d[:64, 0].contiguous(), d[:64, 1].contiguous())
        pa = torch.full((64, 2, 32, 256), 7, dtype=torch.int32, device="cuda")
        pb = pa.clone()
        ta, da = ops.score_frames("DHGR", src, d[1:].contiguous(), table=table, priority=pa)
        tb, db = ops.score_frames("DHGR", src, d[1:].contiguous(), factors=factors, priority=pb)
        return bool(torch.equal(ta, tb) and torch.equal(da, db) and torch.equal(pa, pb))

    ms_table = timed(1.0, table=table)
    ms_table_5pct = timed(0.05, table=table)
    ms = timed(1.0, factors=factors)
    ms_5pct = timed(0.05, factors=factors)
    peaks, peak_src = measured_peaks()
    # one diff_weights bank call = 2 x 32 KiB packed in + 32 KiB int32 out +
    # 8192 x 2 B of table = 112 KiB algorithmic, 8192 x 32 B = 256 KiB of sectors; a DHGR
    # frame is two bank calls.  What a fused launch really streams per frame: 16 KiB screen
    # bytes + 32 KiB source in, 32 KiB packed target + 64 KiB diff out, 64 KiB priorities in
    # and out (the factored kernel reads screen bytes and source once per bank: + 48 KiB).
    alg = 2 * 112 * 1024 * nb
    sector = 2 * 8192 * 32 * nb
    streamed = (16 + 32 + 32 + 64 + 128) * 1024 * nb
    streamed_factored = streamed + (16 + 32) * 1024 * nb
    achieved = alg / (ms * 1e-3) / 1e9
    achieved_table = alg / (ms_table * 1e-3) / 1e9
    # DRAM bytes of one launch of 1024 noise frames from the committed ncu captures
    traffic_table = (1.842528e9 + 0.160930e9) * nb / 1024
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "score_factored_traffic.json")) as f:
            traffic = float(json.load(f)["traffic_bytes_per_launch"]) * nb / 1024
    except (OSError, ValueError, KeyError):
        pass
    return {
        "scored_frames_per_s": nb / (ms * 1e-3),
        "scored_frames_per_s_5pct_change": nb / (ms_5pct * 1e-3),
        "scored_frames_per_s_table_gathers": nb / (ms_table * 1e-3),
        "scored_frames_per_s_table_gathers_5pct_change": nb / (ms_table_5pct * 1e-3),
        "factored_equals_table_path": same_bits(),
        "scored_5pct_note": "same launches on frames that re-draw 5 % of the bytes of their "
                            "predecessor (synth fraction 0.05) instead of all of them: with the "
                            "table, unchanged bytes gather from its diagonal (cache hits); the "
                            "factored kernel does the same work whatever the frame shows",
        "scored_frames_note": (
            "iiv_score_frames_factored: pack + diff_weights (main+aux) + hole mask + priority "
            "fold of %d DISTINCT DHGR frames per launch (source = the previous frame), device "
            "resident; every edit-distance entry evaluated from factor tables in shared memory "
            "(208 KiB per SM, 416 KiB in all, built from the LUT in 20 us), no table in HBM; "
            "%.1f MB streamed per launch (> L2).  scored_frames_per_s_table_gathers: "
            "iiv_score_frames, the same outputs from %d random gathers into the 512 MiB table; "
            "outputs compared in the run" % (nb, streamed_factored / 1e6, 2 * 8192 * nb)),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                     "peak_source": peak_src, "kernel": "score_frames_factored_kernel",
                     "kernel_ms": ms, "algorithmic_bytes_per_launch": alg,
                     "streamed_bytes_per_launch": streamed_factored,
                     "streamed_gbs": streamed_factored / (ms * 1e-3) / 1e9,
                     "lookups_per_s": 2 * 8192 * nb / (ms * 1e-3),
                     "table_gather_kernel": {
                         "kernel": "score_frames_kernel", "kernel_ms": ms_table,
                         "achieved": achieved_table, "frac": achieved_table / peaks["hbm_gbs"],
                         "traffic": traffic_table,
                         "traffic_frac_of_peak": traffic_table / (ms_table * 1e-3) / 1e9
                         / peaks["hbm_gbs"],
                         "gathers_per_s": 2 * 8192 * nb / (ms_table * 1e-3),
                         "bare_gather_ceiling_per_s": 57.8e9,
                         "gather_sector_bytes_per_launch": sector,
                         "streamed_bytes_per_launch": streamed},
                     "note": "algorithmic = 112 KiB per bank call (SURVEY 8(d)) x 2 banks x "
                             "frames, table bytes included although the factored kernel reads "
                             "none.  The table-gather kernel is bound by DRAM row activations "
                             "(93 B of DRAM traffic per 2-byte gather, traffic_frac_of_peak of "
                             "the copy peak, gathers_per_s against the 57.8 G/s of a bare "
                             "gather loop, profiles/r02_gather_flavours.txt); the factored "
                             "kernel replaces a gather by 5 shared-memory loads + ~50 "
                             "instructions and is bound by issue slots and shared-memory bank "
                             "conflicts, with its streams (streamed_gbs) on top"},
    }


def long_clip_figures(torch, ops, table):
    """BASELINE.json configs[3]: the 6000-frame DHGR clip, one launch, its first 600 frames
    checked against the digests of the unmodified reference."""
    from iivision_b200 import synth
    lc = synth.LONG_CLIP
    n = int(os.environ.get("IIV_BENCH_LONG_FRAMES", str(lc["n_frames"])))
    frames = synth.long_clip_frames(n)
    plan = ops.SegmentPlan(synth.movie_schedule("DHGR", n))
    tmem, tpacked = _encode_setup(torch, ops, frames[None])
    ms, all_ms, opc, info, _ = _timed_encode(torch, ops, tmem, tpacked, plan, table,
                                             [lc["rng_seed"]], reps=1, warm=0)
    out = {"frames": n, "opcodes": plan.total, "ms": ms,
           "frames_per_s": n / (ms * 1e-3), "us_per_opcode": ms * 1e3 / plan.total,
           "real_opcodes": int(info.cpu().numpy()[0][:, 0].sum()),
           "note": "one DHGR clip, one thread block, one launch; 980 opcodes per frame, "
                   "bank flip every 292 (Movie.encode schedule)"}
    gold = golden_streams()
    if gold is not None:
        got = opc.cpu().numpy()[0]
        checks = {}
        for frames_done, want in gold["long"]["opcode_sha256"].items():
            k = int(frames_done)
            if k <= n:
                checks[frames_done] = synth.opcode_digest(got[:k * 980]) == want
        out["reference_digest_frames"] = max([int(k) for k in checks] or [0])
        out["matches_reference"] = bool(checks) and all(checks.values())
    return out


def batch_clips_figures(torch, ops, table, world=1, rank=0, dist=None):
    """BASELINE.json configs[4]: 64 independent DHGR clips, clip-per-GPU sharding (8 per
    GPU at N = 8), no collective on the data path; the clips whose reference digests are
    committed are checked on whichever rank holds them."""
    import numpy as np
    from iivision_b200 import parallel, synth
    bc = synth.BATCH_CLIPS
    lo, hi = parallel.shard_range(bc["n_clips"], world, rank)
    n = bc["n_frames"]
    clips = np.stack([synth.batch_clip_frames(c) for c in range(lo, hi)])
    seeds = [synth.batch_clip_seeds(c)[1] for c in range(lo, hi)]
    plan = ops.SegmentPlan(synth.movie_schedule("DHGR", n))
    tmem, tpacked = _encode_setup(torch, ops, clips)
    if dist is not None and world > 1:
        dist.barrier()
    ms, all_ms, opc, info, _ = _timed_encode(torch, ops, tmem, tpacked, plan, table, seeds,
                                             reps=5)
    ok, checked = 1, 0
    gold = golden_streams()
    if gold is not None:
        got = opc.cpu().numpy()
        for key, g in gold["batch"].items():
            c = int(key)
            if lo <= c < hi:
                checked += 1
                if synth.opcode_digest(got[c - lo]) != g["opcode_sha256"][str(n)]:
                    ok = 0
    if dist is not None and world > 1:
        t = torch.tensor([ms, -float(ok), float(checked)], device="cuda", dtype=torch.float64)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, ok, checked = float(tmax[0]), int(-tmax[1].item()), int(tsum[2].item())
    return {"clips": bc["n_clips"], "frames_per_clip": n, "clips_per_gpu": hi - lo,
            "n_gpus": world, "ms": ms,
            "encoded_frames_per_s": bc["n_clips"] * n / (ms * 1e-3),
            "reference_digest_clips": checked,
            "matches_reference": bool(ok) if gold is not None else None,
            "note": "64 distinct DHGR clips x 32 frames (distinct frame and RNG seeds), "
                    "clips sharded contiguously over the GPUs, one thread block per clip; "
                    "slowest rank's median CUDA-event time of 5 launches"}


def sharded_scoring_figures(torch, ops, lut, world, rank, dist):
    """Scored frames/s at N GPUs: frames are independent, so every rank scores its own 1 024
    distinct frames (its own factor tables, no collective); the figure is all frames over the
    slowest rank's time."""
    from iivision_b200 import synth
    nb = int(os.environ.get("IIV_BENCH_SCORE_FRAMES", "1024"))
    factors = ops.score_factors("DHGR", lut)
    fr = synth.synthetic_frames("DHGR", nb + 1, 1.0, seed=1 + 1000 * rank)
    d = torch.from_numpy(fr).cuda()
    src = ops.pack("DHGR", d[:nb, 0].contiguous(), d[:nb, 1].contiguous())
    tgt = d[1:].contiguous()
    prio = torch.zeros((nb, 2, 32, 256), dtype=torch.int32, device="cuda")
    for _ in range(3):
        ops.score_frames("DHGR", src, tgt, priority=prio, factors=factors)
    torch.cuda.synchronize()
    dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 20
    ev[0].record()
    for _ in range(reps):
        ops.score_frames("DHGR", src, tgt, priority=prio, factors=factors)
    ev[1].record()
    torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]) / reps], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"scored_frames_per_s": world * nb / (ms * 1e-3), "frames_per_gpu": nb,
            "ms_slowest_rank": ms, "n_gpus": world, "scaling": "weak",
            "note": "iiv_score_frames_factored on %d distinct DHGR frames per rank (frames "
                    "sharded, no collective); all frames / slowest rank's CUDA-event time" % nb}


def scorer_figures(torch, ops, single=True, world=1, rank=0, dist=None):
    """Second hot path, DHGR NTSC, on DISTINCT synthetic data everywhere: (a) the scoring
    prologue batched over frames, (b) bit-exact encoding of independent clips, one block
    per clip, (c) the named configs[3] / configs[4] workloads with their reference
    digests checked inside the run."""
    import numpy as np
    from iivision_b200 import palette, synth
    lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
    table = ops.table_generate("DHGR", lut, layout=ops.LAYOUT_SYMMETRIC)
    out = {}
    out["config4_batch_clips"] = batch_clips_figures(torch, ops, table, world, rank, dist)
    if not single:
        out["scored_frames"] = sharded_scoring_figures(torch, ops, lut, world, rank, dist)
        return out
    out.update(phase_a_figures(torch, ops, table, lut))

    def encode_run(n_clips, n_frames, reps=5):
        clips = np.stack([synth.synthetic_frames("DHGR", n_frames, 1.0, seed=100 + c)
                          for c in range(n_clips)])
        plan = ops.SegmentPlan(synth.movie_schedule("DHGR", n_frames))
        tmem, tpacked = _encode_setup(torch, ops, clips)
        ms, times, _, info, _ = _timed_encode(torch, ops, tmem, tpacked, plan, table,
                                              [7 + c for c in range(n_clips)], reps)
        return ms, times, _trace(info)

    n_clips, n_frames = int(os.environ.get("IIV_BENCH_CLIPS", "148")), 4
    sampler = ClockSampler(torch.cuda.current_device())
    time.sleep(0.3)
    t_enc0 = time.perf_counter()
    ms, all_ms, trace = encode_run(n_clips, n_frames, reps=25)
    out["encoded_trace_clip0"] = trace
    out["encoded_frames_per_s"] = n_clips * n_frames / (ms * 1e-3)
    out["encoded_note"] = ("%d DISTINCT DHGR clips (own frames, own RNG seeds) x %d frames, one "
                           "block per clip; median of %d runs (ms min/median/max: %.2f / %.2f / "
                           "%.2f)" % (n_clips, n_frames, len(all_ms), all_ms[0], ms, all_ms[-1]))
    out["encoded_clocks"] = sampler.summary(t_enc0, time.perf_counter())
    sampler.stop()
    sm_mhz = out["encoded_clocks"].get("sm_mhz") or 1965.0
    out["encoded_ms_from_sm_cycles"] = trace["slowest_clip_sm_cycles"] / (sm_mhz * 1e3)
    ms2, all2, _ = encode_run(2 * n_clips, n_frames, reps=9)
    out["encoded_frames_per_s_2_waves"] = 2 * n_clips * n_frames / (ms2 * 1e-3)
    out["encoded_2_waves_note"] = "%d distinct clips (two per SM, run in waves): %.2f ms" % (
        2 * n_clips, ms2)
    try:
        out["config3_long_clip"] = long_clip_figures(torch, ops, table)
    except Exception as e:   # noqa: BLE001
        out["config3_long_clip"] = {"error": repr(e)}
    # BASELINE.json configs[0]: HGR NTSC, one 60-frame 280x192 clip (980 opcodes per frame,
    # a single bank), next to the reference's algorithm on the host: oracle/scorer.py keeps
    # the reference's cost structure (whole-array numpy calls per opcode, heapq, Python RNG
    # draws; one core, like the reference)
    try:
        out.update(hgr_clip_and_cpu(torch, ops, table_dhgr=table))
    except Exception as e:   # noqa: BLE001
        out["hgr_clip_error"] = repr(e)
    try:
        out.update(facade_figures(torch))
    except Exception as e:   # noqa: BLE001
        out["facade_error"] = repr(e)
    try:
        out.update(movie_figures(torch, ops, table))
    except Exception as e:   # noqa: BLE001
        out["movie_error"] = repr(e)
    ms1, all1, trace1 = encode_run(1, 16)
    out["single_clip_trace"] = trace1
    out["single_clip_frames_per_s"] = 16 / (ms1 * 1e-3)
    out["single_clip_us_per_opcode"] = ms1 * 1e3 / (16 * 980)
    out["single_clip_note"] = ("one DHGR clip of 16 frames (980 opcodes/frame, bank flip every "
                               "292) on one SM; median of %d runs (ms: %s)" % (
                                   len(all1), ", ".join("%.2f" % x for x in all1)))
    return out


def facade_figures(torch, n_frames=8):
    """The reference-facing call for path 2: video.Video.encode_frame pulled the way
    Movie.encode pulls it (a new generator per frame and per bank flip, one next() per
    audio tick), host arrays in and out -- targets are built from host memory maps, the
    encoder state stays on the device (video.py syncs its host attributes on access), both
    global MT19937 generators are kept in step at every generator's end, opcode tuples are
    read on the host."""
    import contextlib
    import io
    import random
    import numpy as np
    from iivision_b200 import palette, screen, synth, video, video_mode

    class Grabber:
        input_frame_rate = 30

    frames = synth.synthetic_frames("DHGR", n_frames + 1, 1.0, seed=100)
    segs = synth.movie_schedule("DHGR", n_frames + 1)
    random.seed(0)
    np.random.seed(0)
    v = video.Video(Grabber(), 14700., mode=video_mode.VideoMode.DHGR,
                    palette=palette.Palette.NTSC)
    pulled = 0
    t0 = None
    op_seq = None
    h2d = d2h = 0
    tgt, tgt_frame = None, -1
    with contextlib.redirect_stdout(io.StringIO()):
        for frame, is_aux, budget in segs:
            if frame == 1 and t0 is None:      # frame 0 is the warm-up
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            new_target = frame != tgt_frame
            if new_target:                     # one target bitmap per frame (movie.py:81-91)
                tgt_frame = frame
                tgt = screen.DHGRBitmap(
                    palette=palette.Palette.NTSC,
                    main_memory=screen.MemoryMap(1, frames[frame, 0].copy()),
                    aux_memory=screen.MemoryMap(1, frames[frame, 1].copy()))
            op_seq = v.encode_frame(tgt, is_aux=bool(is_aux))
            for _ in range(budget):
                next(op_seq)
            if t0 is not None:
                pulled += budget
                # per generator: opcode records, segment info and the state's tail (both
                # MT19937 states + flags) come down; per frame the target goes up once
                # (two memory maps; the packed words are made on the device).  The encoder
                # state itself stays on the device.
                d2h += budget * 8 + 64 + 2 * 2560 + 64
                if new_target:
                    h2d += 2 * 8192
                    d2h += 32768        # Bitmap.packed of the target, a host attribute
        op_seq.close()
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"facade_frames_per_s": n_frames / dt,
            "facade_us_per_opcode": dt * 1e6 / pulled,
            "facade_h2d_bytes_per_frame": h2d // n_frames,
            "facade_d2h_bytes_per_frame": d2h // n_frames,
            "facade_note": ("video.Video.encode_frame under the Movie.encode schedule, %d DHGR "
                            "frames after one warm-up frame, %d generators, host arrays in and "
                            "out; wall clock" % (n_frames, sum(1 for s_ in segs if s_[0] > 0)))}


def movie_figures(torch, ops, table, n_encoded=16):
    """A whole DHGR movie -- frames and audio samples on the host in, .a2m bytes on the host
    out -- through movie.transcode_device: schedule on the host, one encoder launch, one
    byte-layout launch (30 fps input, every second frame encoded, 14 700 ticks/s as
    main.py's defaults)."""
    import random
    import numpy as np
    from iivision_b200 import movie, opcodes, synth
    n_grabbed = 2 * n_encoded
    frames = synth.synthetic_frames("DHGR", n_grabbed, 1.0, seed=200)
    samples = np.random.default_rng(200).integers(-15, 17, size=490 * n_grabbed - 1)
    try:
        addresses = opcodes.address_table()
    except ValueError:          # no player/iivision.dbg here: any address table times the same
        addresses = (np.arange(1024, dtype=np.uint16).reshape(32, 32) + 0x8000, 0x7f00, 0x7f80)
    pad = np.zeros(640, np.uint32)
    times = []
    n_bytes = 0
    for rep in range(4):
        st = ops.new_clip_states(1)
        pad[:625] = ops.mt_from_python(random.Random(0).getstate())
        ops.state_field(st, ops.F_MT_PY, torch.int32, (640,)).copy_(
            torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
        pad[:625] = ops.mt_from_numpy(np.random.RandomState(0).get_state())
        ops.state_field(st, ops.F_MT_NP, torch.int32, (640,)).copy_(
            torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        data, plan = movie.transcode_device("DHGR", frames, samples, table, st,
                                            every_n_video_frames=2, addresses=addresses)
        host = data.cpu()
        dt = time.perf_counter() - t0
        n_bytes = host.numel()
        if rep:
            times.append(dt)
    times.sort()
    dt = times[len(times) // 2]
    return {"movie_frames_per_s": len(plan.frames_used) / dt,
            "movie_stream_bytes": n_bytes,
            "movie_note": ("movie.transcode_device: %d grabbed DHGR frames (%d encoded), %d "
                           "audio ticks = %.1f s of playback, host arrays in, .a2m bytes on the "
                           "host out; wall clock, median of %d" % (
                               n_grabbed, len(plan.frames_used), plan.emitted,
                               plan.emitted / 14700., len(times)))}


def hgr_clip_and_cpu(torch, ops, table_dhgr):
    import random
    import numpy as np
    from iivision_b200 import palette, synth
    from oracle import scorer
    out = {}
    lut = ops.lut_cie2000(palette.NTSCPalette.rgb_by_value())
    table = ops.table_generate("HGR", lut, layout=ops.LAYOUT_SYMMETRIC)
    n_frames = 60
    frames = synth.synthetic_frames("HGR", n_frames, 1.0, seed=7)
    segs = ops.SegmentPlan(synth.movie_schedule("HGR", n_frames))
    tmem = torch.from_numpy(np.ascontiguousarray(frames[None])).cuda()
    tpacked = ops.pack("HGR", tmem[0, :, 0].contiguous()).view(1, n_frames, 32, 128)
    pad = np.zeros(640, np.uint32)

    def fresh():
        st = ops.new_clip_states(1)
        pad[:625] = ops.mt_from_python(random.Random(0).getstate())
        ops.state_field(st, ops.F_MT_PY, torch.int32, (640,)).copy_(
            torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
        pad[:625] = ops.mt_from_numpy(np.random.RandomState(0).get_state())
        ops.state_field(st, ops.F_MT_NP, torch.int32, (640,)).copy_(
            torch.from_numpy(pad.view(np.int32)).cuda().expand(1, 640))
        return st
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    opc = torch.empty((1, n_frames * 980, 8), dtype=torch.uint8, device="cuda")
    info = torch.zeros((1, len(segs), 8), dtype=torch.int64, device="cuda")
    times = []
    for r in range(4):
        st = fresh()
        torch.cuda.synchronize()
        ev[0].record()
        ops.encode_clips("HGR", st, tmem, tpacked, segs, table, opcodes=opc, seg_info=info)
        ev[1].record()
        torch.cuda.synchronize()
        if r:
            times.append(ev[0].elapsed_time(ev[1]))
    times.sort()
    out["hgr_60_frame_clip_frames_per_s"] = n_frames / (times[len(times) // 2] * 1e-3)
    gpu_ops = opc.cpu().numpy()[0]
    # the same clip's first frames through the CPU restatement of the reference's path
    host_table = table.cpu().numpy()
    cpu_frames = 3
    v = scorer.OracleVideo("HGR", host_table, py_rng=random.Random(0),
                           np_rng=np.random.RandomState(0))
    t0 = time.perf_counter()
    cpu_ops = []
    for fr in range(cpu_frames):
        seq = v.encode_frame(v.target_bitmap(frames[fr, 0]), False)
        for _ in range(980):
            page, content, offs = next(seq)
            cpu_ops.append([page, content] + list(offs))
    dt = time.perf_counter() - t0
    out["hgr_cpu_port_frames_per_s"] = cpu_frames / dt
    out["hgr_cpu_port_note"] = ("oracle/scorer.py (numpy + heapq restatement of screen.py / "
                                "video.py, 1 core) on the first %d frames of the same clip: "
                                "%.0f us per opcode; its opcodes equal the GPU's: %s" % (
                                    cpu_frames, dt * 1e6 / (cpu_frames * 980),
                                    bool(np.array_equal(np.array(cpu_ops, np.uint8),
                                                        gpu_ops[:cpu_frames * 980, :6]))))
    del table
    return out


def all_table_files(torch, dist, make_data_tables, rank, world):
    """make_data_tables.main() into a temporary directory: HGR + DHGR x IIGS + NTSC, each
    generated and deflated on a GPU and written as the reference's compressed .npz.  Ranks
    take whole files (parallel.shard_jobs); wall clock between two barriers, max over ranks;
    one warm-up pass first (page-locked staging, code tables, the first file's cold start)."""
    import shutil
    import tempfile
    tmp = tempfile.mkdtemp(prefix="iiv_tables_%d_" % rank)
    old_dir = make_data_tables.DATA_DIR
    make_data_tables.DATA_DIR = tmp
    try:
        import contextlib
        import io
        sizes = []
        dts = []
        for rep_ in range(2):
            for f in os.listdir(tmp):
                os.unlink(os.path.join(tmp, f))
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                written = make_data_tables.main(rank, world)
            if world > 1:
                dist.barrier()
            dts.append(time.perf_counter() - t0)
            sizes = [os.path.getsize(f) for f in written]
        dt = dts[-1]
        total = sum(sizes)
        n_files = len(sizes)
        if world > 1:
            t = torch.tensor([dt, float(total), float(n_files)], dtype=torch.float64, device="cuda")
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            dt, total, n_files = float(mx[0]), float(t[1]), int(t[2])
        return {"seconds": dt, "first_pass_seconds": dts[0], "files": n_files,
                "bytes": int(total), "raw_bytes": 3 * (1 << 30) + 4 * 128,
                "ranks_with_work": min(world, 4),
                "note": "make_data_tables.main(): the reference's four table files "
                        "(3 GiB of tables) generated + deflated on the device and written; "
                        "files are independent jobs shared out over the ranks (HGR files cost "
                        "twice a DHGR file: 2 ranks take 1 + 1 each, 4 ranks one file each, "
                        "more ranks idle); second pass timed"}
    finally:
        make_data_tables.DATA_DIR = old_dir
        shutil.rmtree(tmp, ignore_errors=True)


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    from iivision_b200 import colours, make_data_tables, ops, palette, parallel, screen

    if args.scorer_only:
        print(json.dumps({"scorer": scorer_figures(torch, ops)}))
        return
    # one process per GPU: keep the process (and every page it first touches or page-locks)
    # on the NUMA node its GPU hangs off
    numa_node = parallel.bind_to_gpu_numa_node(local_rank) if world > 1 else None
    pal = palette.NTSCPalette
    edp = make_data_tables.compute_substitute_costs(pal)       # LUT on the device (FP64)
    lut = make_data_tables._lut16(edp.substitute_costs)
    table = torch.empty(ops.table_shape(MODE), dtype=torch.uint16, device="cuda")
    stream = torch.cuda.current_stream()

    def sharded_step(how):
        if how == "nccl":
            return parallel.generate_sharded(MODE, lut, layout=ops.LAYOUT_SYMMETRIC, out=xtable)
        return parallel.generate_sharded_fused(MODE, lut, layout=ops.LAYOUT_SYMMETRIC,
                                               multicast=(how == "fused_mc"))

    def step():
        # N = 1: the table.  N > 1: the table on EVERY GPU, which is what the scorer needs
        # (each clip's encoder gathers from its own GPU's copy): every rank generates its
        # own replica.  One GPU makes the table (0.26 ms) faster than NVLink can deliver
        # (N-1)/N of it (>= 1.2 ms at 8 GPUs), so replicas beat any exchange; the sharded
        # variants are timed below as alt_exchange.
        ops.table_generate(MODE, lut, layout=ops.LAYOUT_SYMMETRIC, out=table)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allmin_int(x):
        if world == 1:
            return int(x)
        t = torch.tensor([int(x)], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.3 if rank == 0 else 0)
    # warm-up: at least W (>= 3) steps AND >= 1.2 s of back-to-back generator
    # launches, so that clocks settle under load and the 100 ms clock sampler gets
    # readings under the same load as the timed region (which lasts only K x ~0.3 ms)
    t_load0 = time.perf_counter()
    done = 0
    while True:
        for _ in range(8):
            step()
        torch.cuda.synchronize()
        done += 8
        more = done < max(args.warmup, 3) or time.perf_counter() - t_load0 < 1.2
        if world > 1:   # same iteration count on every rank
            flag = torch.tensor([int(more)], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            more = bool(flag.item())
        if not more:
            break
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_host0 = time.perf_counter()
    evs[0].record(stream)
    for k in range(args.steps):
        step()
        evs[k + 1].record(stream)
    barrier()
    t_host1 = time.perf_counter()
    per_step = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    total_ms = allmax(evs[0].elapsed_time(evs[-1]))
    # every rank delivered one whole table per step
    value = world * ENTRIES * args.steps / (total_ms * 1e-3)

    # outside the timed region: the replicas are the same table everywhere
    replicas_identical = None
    if world > 1:
        sig = torch.stack([table.view(torch.int32).sum(dtype=torch.int64),
                           table.view(torch.int32)[::4097].sum(dtype=torch.int64)])
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        replicas_identical = all(bool(torch.equal(x, sigs[0])) for x in sigs)

    alt = None
    if world > 1:
        # sharded generation + exchange (rows split over the ranks, every rank ends up with
        # the whole table), same K and barriers; each result is compared with this rank's
        # own replica outside the timed region
        alt = []
        xtable = torch.empty_like(table)
        for how in ("fused_peer", "fused_mc", "nccl"):
            ok = 1
            res = None
            try:
                for _ in range(3):
                    res = sharded_step(how)
                torch.cuda.synchronize()
            except Exception as e:          # e.g. no multicast on this fabric  # noqa: BLE001
                ok = 0
                err = str(e).splitlines()[0][:80]
            if not allmin_int(ok):
                alt.append({"exchange": how, "unavailable": err if not ok else "on another rank"})
                continue
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(args.steps):
                res = sharded_step(how)
            a1.record(stream)
            barrier()
            ms = allmax(a0.elapsed_time(a1)) / args.steps
            same = allmin_int(int(torch.equal(res.view(torch.int16), table.view(torch.int16))))
            alt.append({"exchange": how, "ms_per_step": ms,
                        "tables_per_s_whole_job": 1e3 / ms,
                        "entries_per_s_one_table": ENTRIES / (ms * 1e-3),
                        "equals_local_replica_on_every_rank": bool(same)})
        del xtable

    # kernel-only timing for the roofline: the generator kernel of this rank's table
    kev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kernel_ms = []
    for _ in range(min(args.steps, 10)):
        kev[0].record(stream)
        ops.table_generate(MODE, lut, layout=ops.LAYOUT_SYMMETRIC, out=table)
        kev[1].record(stream)
        torch.cuda.synchronize()
        kernel_ms.append(kev[0].elapsed_time(kev[1]))
    # reference points: write-only streams over the same 1 GiB buffer.  cudaMemsetAsync and
    # a one-store-per-thread fill reach 7.2-7.4 TB/s on this part (profiles/r02_hbm_fill.txt)
    # -- above the read+write copy figure MEASURED_PEAKS.json records; torch's fill_ kernel
    # does not (3.9 TB/s) and is no ceiling.
    fill_gbs = {}
    for name, variant in (("memset", 0), ("one_store_per_thread", 1)):
        # timed the way the steps are: K launches back to back between two events
        for _ in range(3):
            ops.fill_probe(table, variant)
        torch.cuda.synchronize()
        kev[0].record(stream)
        for _ in range(args.steps):
            ops.fill_probe(table, variant)
        kev[1].record(stream)
        torch.cuda.synchronize()
        fill_gbs[name] = 2.0 * ENTRIES * args.steps / (kev[0].elapsed_time(kev[1]) * 1e-3) / 1e9
    ops.table_generate(MODE, lut, layout=ops.LAYOUT_SYMMETRIC, out=table)
    clocks = sampler.summary(t_load0, time.perf_counter()) if sampler else None
    # the generator's average duration over the TIMED REGION (CUDA events on its stream
    # around the K back-to-back steps of this rank; a step = split_prologue, ~9 us,
    # + split_kernel launched as its programmatic dependent, so `achieved` charges the
    # dominant kernel with the prologue).  The same step timed alone
    # with a synchronize on both sides (kernel_ms_synced) is ~10 % longer: on an idle GPU the
    # three launches wait for the host to submit them.  The write-only fills below are timed
    # back to back like the steps, so frac_of_write_only compares like with like.
    k_ms_synced = sum(kernel_ms) / len(kernel_ms)
    k_ms = evs[0].elapsed_time(evs[-1]) / args.steps
    peaks, peak_src = measured_peaks()
    alg_bytes = 2.0 * ENTRIES
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    achieved_synced = alg_bytes / (k_ms_synced * 1e-3) / 1e9

    # e2e through the reference-facing call: host parameters in, host numpy array out
    e2e_steps = max(2, min(args.steps, 5))
    # bytes one rank's banded lower-triangle download moves (ops.table_download)
    d2h_bytes = 0
    for b, e in parallel.triangle_row_blocks(1 << BITS, world, rank):
        per = -(-(e - b) // ops.DOWNLOAD_BANDS)
        for r0 in range(b, e, per):
            r1 = min(r0 + per, e)
            d2h_bytes += N_OFF * (r1 - r0) * min((2 * (r1 - 1) + 63) // 64 * 64, 2 << BITS)
    host = None
    if world == 1:
        def e2e_step():
            edp_ = make_data_tables.compute_substitute_costs(pal)
            return make_data_tables.compute_edit_distance(
                edp_, screen.HGRBitmap, colours.HGRColours)
    else:
        host = parallel.SharedHostTable(MODE)

        def e2e_step():
            # every rank: LUT (device FP64), its row block, its own D2H into its page-locked
            # slice of the one shared host array; rank 0's caller holds the array
            edp_ = make_data_tables.compute_substitute_costs(pal)
            return parallel.compute_edit_distance_sharded(
                MODE, make_data_tables._lut16(edp_.substitute_costs), host,
                device_table=table)
    e2e_step()
    barrier()
    res = None
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = None          # the caller drops one table before asking for the next
        res = e2e_step()
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_verified = None
    if world > 1:
        # rank 0 checks the assembled host array against a table it generates alone
        if rank == 0:
            ops.table_generate(MODE, lut, layout=ops.LAYOUT_TRIANGULAR, out=table)
            e2e_verified = bool(np.array_equal(
                res, table.cpu().view(torch.int16).numpy().view(np.uint16)))
        barrier()
    # the e2e step is the device-to-host copy of the table: time that copy alone, same
    # page-locked destination, so the figure can be read against this box's PCIe rate
    d2h_ms = d2h_full_ms = None
    if world == 1 and res is not None:
        del res
        res = None
        # the bare copies next to it: the banded lower-triangle download the call uses, and
        # a whole-table copy (what moving the zeros too would cost) -- same page-locked pool
        lease, arr = make_data_tables._pinned_pool.lease(tuple(table.shape))
        ops.table_generate(MODE, lut, layout=ops.LAYOUT_TRIANGULAR, out=table)
        evc = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        evc[0].record()
        ops.table_download(MODE, table, lease.data_ptr(), layout=ops.LAYOUT_TRIANGULAR)
        evc[1].record()
        torch.cuda.synchronize()
        d2h_ms = evc[0].elapsed_time(evc[1])
        evc[0].record()
        ops.table_download(MODE, table, lease.data_ptr(), layout=ops.LAYOUT_SYMMETRIC)
        evc[1].record()
        torch.cuda.synchronize()
        d2h_full_ms = evc[0].elapsed_time(evc[1])
        del arr, lease
    del res
    if host is not None:
        host.close()
    if sampler:
        sampler.stop()

    # second hot path at N > 1: BASELINE.json configs[4], 64 distinct clips sharded
    # clip-per-GPU with no collective on the data path (every rank takes part)
    clips_multi = None
    if world > 1 and not args.no_scorer:
        try:
            clips_multi = scorer_figures(torch, ops, single=False, world=world, rank=rank,
                                         dist=dist)
        except Exception as e:   # noqa: BLE001
            clips_multi = {"error": repr(e)}
    nodes = None
    if world > 1:
        nodes = [None] * world
        dist.all_gather_object(nodes, numa_node)
    # the table job a user waits for: make_data_tables.main(), the four .npz files of the
    # reference (make_data_tables.py:191-204), as independent jobs over the ranks
    table_files = None
    if not args.no_scorer:
        try:
            table_files = all_table_files(torch, dist, make_data_tables, rank, world)
        except Exception as e:   # noqa: BLE001
            table_files = {"error": repr(e)}
    if rank != 0:
        return
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3, done),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16",
        "data": "synthetic (palette constants; no external data)",
        "config": workload_config(world),
        "clocks": clocks,
        "e2e": {"value": ENTRIES * e2e_steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": 16 * 3 + 256 * 4,
                "d2h_bytes_per_step": d2h_bytes + 256 * 4,
                "steps": e2e_steps,
                **({"d2h_copy_ms": d2h_ms, "d2h_whole_table_copy_ms": d2h_full_ms,
                    "d2h_whole_table_copy_gbs": 2.0 * ENTRIES / (d2h_full_ms * 1e-3) / 1e9}
                   if d2h_ms else {}),
                **({"assembled_array_equals_single_gpu_table": e2e_verified,
                    "gpu_numa_nodes": nodes} if world > 1 else {}),
                "call": "make_data_tables.compute_substitute_costs + compute_edit_distance "
                        "-> host uint16 array (only j < i crosses PCIe: the array's other "
                        "half is zero by construction)" if world == 1 else
                        "compute_substitute_costs + parallel.compute_edit_distance_sharded: ONE "
                        "host uint16 array (the reference's lower-triangular table) in shared "
                        "page-locked memory; each rank generates its row block and copies it "
                        "over its own PCIe link; d2h bytes are per rank"},
        "gpu_launches": ops.launches_per_table_generate() * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": ncu_traffic(),
                     "peak_source": peak_src, "kernel": ops.generator_kernel_name(),
                     "kernel_ms": k_ms,
                     "kernel_ms_synced": k_ms_synced,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "write_only_gbs": fill_gbs,
                     "frac_of_write_only": achieved / max(fill_gbs.values()),
                     "note": "2 B stored per entry x entries per launch / average duration "
                             "of a step over the timed region (CUDA events on the launching "
                             "stream around the K steps; split_kernel is 94 % of a step, its "
                             "prologue is charged to it); kernel_ms_synced = the same step "
                             "timed alone between synchronizes (host submission exposed); peak = measured "
                             "copy (read+write) bandwidth, which a write-only stream can exceed: "
                             "write_only_gbs = cudaMemsetAsync and a one-16-byte-store-per-thread "
                             "kernel over the same 1 GiB timed in this run back to back like the steps, the ceiling that "
                             "applies (profiles/r02_hbm_fill.txt) -- the generator stores every "
                             "byte exactly once and reads 4 MiB of its own tables from L2"},
        "step_ms_min_max": [min(per_step), max(per_step)],
        "wall_s_timed_region": t_host1 - t_host0,
    }
    if table_files is not None:
        line["table_files"] = table_files
    if world > 1:
        line["replicas_identical"] = replicas_identical
        line["alt_exchange"] = alt
    if world == 1:
        # the reference's whole job (README "about 90 minutes"): both modes x both palettes
        try:
            jobs = [(m, make_data_tables._lut16(
                        make_data_tables.compute_substitute_costs(p).substitute_costs))
                    for p in (palette.IIGSPalette, palette.NTSCPalette) for m in ("HGR", "DHGR")]
            outs = {m: torch.empty(ops.table_shape(m), dtype=torch.uint16, device="cuda")
                    for m in ("HGR", "DHGR")}
            for m, l in jobs:
                ops.table_generate(m, l, layout=ops.LAYOUT_TRIANGULAR, out=outs[m])
            torch.cuda.synchronize()
            kev[0].record(stream)
            for m, l in jobs:
                ops.table_generate(m, l, layout=ops.LAYOUT_TRIANGULAR, out=outs[m])
            kev[1].record(stream)
            torch.cuda.synchronize()
            line["all_four_tables"] = {
                "ms": kev[0].elapsed_time(kev[1]),
                "note": "HGR + DHGR x IIGS + NTSC in the reference's file layout (j < i), "
                        "3 GiB, 805 240 832 dam_lev evaluations; device time, tables left in HBM"}
            del outs
        except Exception as e:
            line["all_four_tables"] = {"error": repr(e)}
    if world == 1 and not args.no_scorer:
        # make_edit_distance end to end: LUT + table + the reference's compressed .npz
        # (make_data_tables.py:177-188); the deflate runs on the device (csrc/iiv_deflate.cu)
        # and only the compressed stream crosses PCIe
        try:
            import tempfile
            with tempfile.TemporaryDirectory() as tmp:
                old_dir = make_data_tables.DATA_DIR
                make_data_tables.DATA_DIR = tmp
                try:
                    t0 = time.perf_counter()
                    make_data_tables.make_edit_distance(
                        pal, make_data_tables.compute_substitute_costs(pal),
                        screen.HGRBitmap, colours.HGRColours)
                    dt = time.perf_counter() - t0
                finally:
                    make_data_tables.DATA_DIR = old_dir
                path = os.path.join(tmp, "HGR_palette_%d_edit_distance.npz" % pal.ID.value)
                line["npz_file"] = {
                    "seconds": dt, "bytes": os.path.getsize(path),
                    "raw_bytes": 128 + (2 << 28) * 2,
                    "note": "make_data_tables.make_edit_distance(HGR, NTSC) into a temporary "
                            "directory, first call of the process: the file the reference's "
                            "loader reads (np.load(...)['edit_distance']); table generated and "
                            "deflated on the device (dynamic-Huffman blocks of 32 KiB), the "
                            "compressed stream staged home through a page-locked ring while "
                            "the file is written; zlib -6 on 16 host threads makes 295.5 MB "
                            "in 1.3 s (profiles/r02c_deflate_hgr.txt)"}
        except Exception as e:   # noqa: BLE001
            line["npz_file"] = {"error": repr(e)}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    if world == 1 and not args.no_scorer:
        try:
            line["scorer"] = scorer_figures(torch, ops)
        except Exception as e:  # secondary figures must not lose the headline
            line["scorer"] = {"error": repr(e)}
    if world > 1 and clips_multi is not None:
        line["scorer"] = clips_multi
    print(json.dumps(line))


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
