"""Shared helpers: run the same clip through the numpy oracle and the CUDA path."""

import random

import numpy as np


def run_oracle(mode, table, frames, segments, seed):
    """Returns (opcodes int[n,6], real flags, OracleVideo, py_rng, np_rng)."""
    from oracle import scorer
    py = random.Random(seed)
    npr = np.random.RandomState(seed)
    v = scorer.OracleVideo(mode, table, py_rng=py, np_rng=npr)
    ops = []
    real = []
    for frame, is_aux, budget in segments:
        if budget == 0:
            continue
        tgt = v.target_bitmap(frames[frame, 0],
                              frames[frame, 1] if mode == "DHGR" else None)
        v.out_of_work = {True: False, False: False}
        seq = v.encode_frame(tgt, bool(is_aux))
        for _ in range(budget):
            page, content, offs = next(seq)
            real.append(0 if v.out_of_work[bool(is_aux)] else 1)
            ops.append([int(page), int(content)] + [int(o) for o in offs])
    return np.array(ops, dtype=np.int64).reshape(-1, 6), np.array(real), v, py, npr


def seed_states(ops_mod, states, seeds):
    """Video.__init__ + random.seed(s); np.random.seed(s) for each clip."""
    import torch
    n = states.shape[0]
    mt_np = np.zeros((n, 640), dtype=np.uint32)
    mt_py = np.zeros((n, 640), dtype=np.uint32)
    for k, s in enumerate(seeds):
        mt_py[k, :625] = ops_mod.mt_from_python(random.Random(s).getstate())
        mt_np[k, :625] = ops_mod.mt_from_numpy(np.random.RandomState(s).get_state())
    ops_mod.state_field(states, ops_mod.F_MT_NP, torch.int32, (640,)).copy_(
        torch.from_numpy(mt_np.view(np.int32)).cuda())
    ops_mod.state_field(states, ops_mod.F_MT_PY, torch.int32, (640,)).copy_(
        torch.from_numpy(mt_py.view(np.int32)).cuda())


def run_device(ops_mod, mode, dtable, frames_by_clip, segments, seeds):
    """frames_by_clip uint8[n_clips, n_frames, banks, 32, 256]."""
    import torch
    n_clips = frames_by_clip.shape[0]
    states = ops_mod.new_clip_states(n_clips)
    seed_states(ops_mod, states, seeds)
    tmem = torch.from_numpy(np.ascontiguousarray(frames_by_clip)).cuda()
    flat = tmem.view(-1, tmem.shape[2], 32, 256)
    main = flat[:, 0].contiguous()
    aux = flat[:, 1].contiguous() if mode == "DHGR" else None
    tpacked = ops_mod.pack(mode, main, aux).view(n_clips, -1, 32, 128)
    opcodes, info = ops_mod.encode_clips(mode, states, tmem, tpacked, segments, dtable)
    torch.cuda.synchronize()
    return opcodes.cpu().numpy(), info.cpu().numpy(), states
