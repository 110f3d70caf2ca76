"""Where a video.Video generator's wall time goes under the Movie.encode schedule."""
import contextlib, io, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from iivision_b200 import palette, screen, synth, video, video_mode

acc = {}
def timed(cls, name):
    fn = getattr(cls, name)
    def wrap(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
            acc[name + "#"] = acc.get(name + "#", 0) + 1
    setattr(cls, name, wrap)
cyc = []
_wait = video._Kernel.wait
def wait2(self, off, events):
    fresh = self._ops is None
    r = _wait(self, off, events)
    if fresh:
        info = self.stage.info.numpy()
        cyc.append((int(info[4]), int(info[5]), self.budget))
    return r
video._Kernel.wait = wait2
for cls, names in ((video._Run, ["__init__", "finish", "_speculate_next", "_rerun"]),
                   (video._Kernel, ["wait"]), (video.Video, ["_launch", "_target"]),
                   (screen.Bitmap, ["_pack"])):
    for n in names:
        timed(cls, n)

class Grabber:
    input_frame_rate = 30
n_frames = 24
frames = synth.synthetic_frames("DHGR", n_frames + 1, 1.0, seed=100)
segs = synth.movie_schedule("DHGR", n_frames + 1)
random.seed(0); np.random.seed(0)
v = video.Video(Grabber(), 14700., mode=video_mode.VideoMode.DHGR, palette=palette.Palette.NTSC)
tgt, tgt_frame, t0, pulled, loop = None, -1, None, 0, 0.0
with contextlib.redirect_stdout(io.StringIO()):
    for frame, is_aux, budget in segs:
        if frame == 1 and t0 is None:
            torch.cuda.synchronize(); acc.clear(); t0 = time.perf_counter()
        if frame != tgt_frame:
            tgt_frame = frame
            tgt = screen.DHGRBitmap(palette=palette.Palette.NTSC,
                                    main_memory=screen.MemoryMap(1, frames[frame, 0].copy()),
                                    aux_memory=screen.MemoryMap(1, frames[frame, 1].copy()))
        op_seq = v.encode_frame(tgt, is_aux=bool(is_aux))
        t1 = time.perf_counter()
        for _ in range(budget):
            next(op_seq)
        loop += time.perf_counter() - t1
        if t0 is not None: pulled += budget
    op_seq.close(); torch.cuda.synchronize()
dt = time.perf_counter() - t0
gens = sum(1 for s in segs if s[0] > 0)
print("%d frames, %d generators: %.1f frames/s, %.1f us per generator" % (n_frames, gens, n_frames / dt, dt * 1e6 / gens))
print("next() loops incl. first-next work: %.1f us per generator" % (loop * 1e6 / len(segs)))
for k in sorted(k for k in acc if not k.endswith("#")):
    print("  %-18s %8.1f us per call x %d" % (k, acc[k] * 1e6 / acc[k + "#"], acc[k + "#"]))
c = np.array(cyc[-100:], dtype=np.float64)
print("device cycles per generator (last 100): score+heapify %.0f, opcode loop %.0f (%.0f per opcode); at 1965 MHz: %.1f + %.1f us" % (
    c[:, 0].mean(), c[:, 1].mean(), c[:, 1].sum() / c[:, 2].sum(), c[:, 0].mean() / 1965, c[:, 1].mean() / 1965))
