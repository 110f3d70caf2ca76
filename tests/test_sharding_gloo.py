"""CPU, world_size 2 over gloo: the multi-GPU host logic (row-block sharding +
in-place all-gather of the table, clip-per-rank sharding).  The CUDA generator
is replaced by the oracle here, purely so the exchange can run without a GPU."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_generate(mode, lut, layout, begin, end, out):
    from oracle import tables
    name = {0: "HGR", 1: "DHGR"}[mode]
    view = out.numpy().view(np.uint16) if out.dtype != torch.uint16 else out.view(torch.int16).numpy().view(np.uint16)
    tables.build_table(name, lut, begin, end, triangular=(layout == 0), out=view)


def _host_download(mode, table, host_ptr, begin, end):
    """Stand-in for ops.table_download on CPU tensors: the rows' lower-triangle columns."""
    import ctypes
    n = 1 << 13
    src = table.view(torch.int16).numpy().view(np.uint16).reshape(4, n, n)
    dst = np.ctypeslib.as_array(ctypes.cast(host_ptr, ctypes.POINTER(ctypes.c_uint16)),
                                shape=(4, n, n))
    width = max(end - 1, 0)
    dst[:, begin:end, :width] = src[:, begin:end, :width]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from iivision_b200 import parallel
        from oracle import tables
        lut = tables.substitution_lut(5)
        for layout in (1, 0):
            out = torch.zeros((4, 1 << 26), dtype=torch.int16).view(torch.uint16)
            parallel.generate_sharded("DHGR", lut, layout=layout, out=out,
                                      generate_fn=_oracle_generate)
            want, _ = tables.build_table("DHGR", lut, triangular=(layout == 0))
            got = out.view(torch.int16).numpy().view(np.uint16)
            assert np.array_equal(got, want), "layout %d rank %d" % (layout, rank)
        # host delivery: every rank copies its row block into one shared host array
        host = parallel.SharedHostTable("DHGR", register=False)
        try:
            dev = torch.zeros((4, 1 << 26), dtype=torch.int16).view(torch.uint16)
            got = parallel.compute_edit_distance_sharded(
                "DHGR", lut, host, device_table=dev, generate_fn=_oracle_generate,
                download_fn=_host_download)
            want, _ = tables.build_table("DHGR", lut, triangular=True)
            assert np.array_equal(got, want), "shared host table, rank %d" % rank
        finally:
            host.close()
        lo, hi = parallel.shard_range(5, world, rank)
        got = parallel.gather_clip_outputs(np.arange(lo, hi), 5)
        if rank == 0:
            assert np.concatenate(got).tolist() == [0, 1, 2, 3, 4]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_partitions():
    from iivision_b200 import parallel
    for world in (1, 2, 4, 8):
        parts = parallel.row_partition(8192, world)
        assert parts[0][0] == 0 and parts[-1][1] == 8192
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert len({e - b for b, e in parts}) == 1
    with pytest.raises(ValueError):
        parallel.row_partition(8192, 3)
    for world in (1, 2, 4, 8):
        blocks = [parallel.triangle_row_blocks(8192, world, r) for r in range(world)]
        flat = sorted(b for bl in blocks for b in bl)
        assert flat[0][0] == 0 and flat[-1][1] == 8192
        assert all(a[1] == b[0] for a, b in zip(flat, flat[1:]))
        below = [sum(sum(range(b, e)) for b, e in bl) for bl in blocks]   # entries j < i
        assert max(below) - min(below) <= 8192 * (8192 // (2 * world))
    for n, world in ((64, 8), (5, 2), (3, 4), (0, 2)):
        spans = [parallel.shard_range(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


@pytest.mark.timeout(300)
def test_sharded_table_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=280) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_jobs_balances_and_covers():
    """Table files are independent jobs: every job goes to exactly one rank, every rank
    computes the same assignment, and the heaviest rank carries no more than it must."""
    from iivision_b200 import parallel
    costs = [2.0, 1.0, 2.0, 1.0]          # HGR / DHGR x two palettes (make_data_tables.main)
    for world in (1, 2, 3, 4, 8):
        shares = [parallel.shard_jobs(costs, world, r) for r in range(world)]
        assert sorted(k for s in shares for k in s) == list(range(len(costs)))
        loads = [sum(costs[k] for k in s) for s in shares]
        assert max(loads) == {1: 6.0, 2: 3.0, 3: 2.0, 4: 2.0, 8: 2.0}[world]
    import pytest
    with pytest.raises(ValueError):
        parallel.shard_jobs(costs, 2, 2)
