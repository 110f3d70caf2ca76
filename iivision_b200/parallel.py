"""Multi-GPU host logic: one process per GPU (torch.distributed), no ML vocabulary.

Two partitions, as the paths allow (DESIGN.md "Multi-GPU"):

* edit-distance tables shard by ROW BLOCKS of the source index ``i``: rank r
  fills rows [begin_r, end_r) of every byte offset directly in its slot of the
  full table, then the shards are exchanged so every GPU holds the whole table
  (that is what the scorer needs).  Exchange = in-place NCCL all-gather per
  offset slice, or the fused generate+scatter kernel that stores each block
  straight into every peer's table over NVLink (``generate_sharded_fused``).
* independent clips shard clip-per-GPU with no data-path collective; a single
  clip's frame sequence is sequential encoder state and stays on one GPU.

The reference has no counterpart (single process, make_data_tables.py:143-172 is
one Python loop); the functions are generic over the buffer/generator so that
the world_size-2 gloo tests on CPU can drive the same code.
"""

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

MASKED_BITS = {_lib.MODE_HGR: 14, _lib.MODE_DHGR: 13}
NUM_OFFSETS = {_lib.MODE_HGR: 2, _lib.MODE_DHGR: 4}
_MODES = {"HGR": _lib.MODE_HGR, "DHGR": _lib.MODE_DHGR}


def _mode_id(mode) -> int:
    return _MODES[mode] if isinstance(mode, str) else int(mode)


def row_partition(n_rows: int, world: int, layout: int = _lib.LAYOUT_SYMMETRIC
                  ) -> List[Tuple[int, int]]:
    """Contiguous row blocks, one per rank, equal in size (NCCL all-gather needs
    equal counts).  n_rows is a power of two and world one of 1, 2, 4, 8."""
    if world < 1 or n_rows % world:
        raise ValueError("world size %d does not divide %d rows" % (world, n_rows))
    step = n_rows // world
    return [(r * step, (r + 1) * step) for r in range(world)]


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous split of independent items (clips) across ranks."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def generate_sharded(mode, lut, layout: int = _lib.LAYOUT_SYMMETRIC,
                     out: Optional[torch.Tensor] = None, group=None,
                     generate_fn: Optional[Callable] = None) -> torch.Tensor:
    """compute_edit_distance sharded by row blocks + in-place all-gather.

    ``generate_fn(mode, lut, layout, row_begin, row_end, out)`` fills the rows
    of ``out`` (uint16[n_off, 4**bits]); default is the CUDA generator.
    Returns the full table on every rank.
    """
    m = _mode_id(mode)
    bits, n_off = MASKED_BITS[m], NUM_OFFSETS[m]
    n_rows = 1 << bits
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if generate_fn is None:
        from . import ops
        generate_fn = ops.table_generate_into
    if out is None:
        out = torch.empty((n_off, 1 << (2 * bits)), dtype=torch.uint16,
                          device="cuda")
    begin, end = row_partition(n_rows, world, layout)[rank]
    generate_fn(m, lut, layout, begin, end, out)
    if world > 1:
        # [o][i][j]: rank r's rows of offset o are one contiguous run, and runs of
        # successive ranks are adjacent -> NCCL's in-place all-gather layout.
        # uint16 is moved as raw bytes (a dtype both NCCL and gloo carry).
        flat = out.view(torch.uint8).view(n_off, n_rows << (bits + 1))
        chunk = (end - begin) << (bits + 1)
        for o in range(n_off):
            dist.all_gather_into_tensor(
                flat[o], flat[o, rank * chunk:(rank + 1) * chunk], group=group)
    return out


_symm_cache = {}


def symmetric_table(mode, group=None):
    """(table, handle): a uint16[n_off, 4**bits] table in NVLink peer-mapped
    (symmetric) memory, allocated and rendezvous-ed once per (mode, group)."""
    import torch.distributed._symmetric_memory as symm
    m = _mode_id(mode)
    group = group if group is not None else dist.group.WORLD
    key = (m, group.group_name)
    if key not in _symm_cache:
        bits, n_off = MASKED_BITS[m], NUM_OFFSETS[m]
        buf = symm.empty((n_off, 1 << (2 * bits)), dtype=torch.int16,
                         device=torch.device("cuda", torch.cuda.current_device()))
        hdl = symm.rendezvous(buf, group)
        _symm_cache[key] = (buf.view(torch.uint16), hdl)
    return _symm_cache[key]


def generate_sharded_fused(mode, lut, layout: int = _lib.LAYOUT_SYMMETRIC, group=None,
                           multicast: Optional[bool] = None) -> torch.Tensor:
    """compute_edit_distance sharded by row blocks with the exchange fused into
    the generator: every rank's kernel stores each finished 16-byte run straight
    into ALL ranks' tables over NVLink (peer-mapped pointers), or once through the
    NVSwitch multicast mapping (multimem.st), so no separate collective runs.
    Returns this rank's full table (symmetric-memory backed, reused across calls).
    """
    from . import ops
    m = _mode_id(mode)
    table, hdl = symmetric_table(m, group)
    world, rank = hdl.world_size, hdl.rank
    begin, end = row_partition(1 << MASKED_BITS[m], world, layout)[rank]
    mc = 0
    if multicast is None or multicast:
        mc = int(hdl.multicast_ptr or 0)      # 0 when the fabric has no multicast
        if multicast and not mc:
            raise RuntimeError("NVSwitch multicast is not available for this allocation")
    hdl.barrier(channel=0)        # peers have finished reading the previous contents
    ops.table_generate_scatter(m, lut, [int(p) for p in hdl.buffer_ptrs], rank, begin, end,
                               layout=layout, multicast_ptr=mc)
    hdl.barrier(channel=1)        # every rank's stores have landed everywhere
    return table


def gather_clip_outputs(local: np.ndarray, n_clips: int, group=None) -> Optional[List[np.ndarray]]:
    """Host-side gather of per-clip opcode buffers to rank 0 (not on the timed
    data path: clips are independent)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [local]
    world = dist.get_world_size(group)
    bucket = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    return bucket
