"""GPU, >= 2 devices: the sharded table generator with its exchange over NCCL
and over NVLink peer stores / NVSwitch multicast, against the single-GPU table."""

import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from iivision_b200 import ops, parallel
        from oracle import tables
        lut = tables.substitution_lut(5)
        report = {}
        for mode in ("DHGR", "HGR"):
            for layout in (ops.LAYOUT_SYMMETRIC, ops.LAYOUT_TRIANGULAR):
                want = ops.table_generate(mode, lut, layout=layout)
                got = parallel.generate_sharded(mode, lut, layout=layout)
                torch.cuda.synchronize()
                assert torch.equal(got.view(torch.int16), want.view(torch.int16)), \
                    "nccl %s %d" % (mode, layout)
                for mc in (False, None):
                    tab = parallel.generate_sharded_fused(mode, lut, layout=layout, multicast=mc)
                    torch.cuda.synchronize()
                    assert torch.equal(tab.view(torch.int16), want.view(torch.int16)), \
                        "fused %s %d mc=%r" % (mode, layout, mc)
                    tab.zero_()
                    torch.cuda.synchronize()
                    dist.barrier()
            _, hdl = parallel.symmetric_table(mode)
            report[mode] = bool(hdl.multicast_ptr)
        q.put((rank, "ok", report))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc(), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.timeout(600)
def test_sharded_generate_nccl_and_fused():
    import torch.multiprocessing as mp
    world = min(_n_gpus(), 8)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=560) for _ in procs]
    for p in procs:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res
    print("multicast used:", res[0][2])
