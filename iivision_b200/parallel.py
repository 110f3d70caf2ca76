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


def triangle_row_blocks(n_rows: int, world: int, rank: int) -> List[Tuple[int, int]]:
    """Row blocks of one rank when only j < i matters (the reference's file layout): the
    rows are cut into 2 * world equal blocks and rank r takes block r and block
    2 * world - 1 - r, so every rank owns the same number of entries below the diagonal
    (to within one block's worth of rows)."""
    if world < 1 or n_rows % (2 * world):
        raise ValueError("2 x world size %d does not divide %d rows" % (world, n_rows))
    step = n_rows // (2 * world)
    lo, hi = rank, 2 * world - 1 - rank
    return [(lo * step, (lo + 1) * step), (hi * step, (hi + 1) * step)]


def shard_jobs(costs, world: int, rank: int) -> List[int]:
    """Indices of the independent jobs (table files) of one rank: longest job first, each to
    the rank with the least work so far (ties: the lower rank).  Every rank computes the
    same assignment, so no message is needed."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d of %d" % (rank, world))
    load = [0.0] * world
    mine = []
    for k in sorted(range(len(costs)), key=lambda k: (-float(costs[k]), k)):
        r = min(range(world), key=lambda r: (load[r], r))
        load[r] += float(costs[k])
        if r == rank:
            mine.append(k)
    return sorted(mine)


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


# ---- host delivery of a sharded table: every GPU copies its rows home itself -------------
#
# compute_edit_distance returns a HOST array (make_data_tables.py:111-174).  On one GPU the
# whole call is the 1 GiB device-to-host copy (~19 ms at 55 GB/s against 0.26 ms of kernel).
# That copy shards exactly like the rows do: rank r generates rows [begin_r, end_r) and
# copies them over ITS OWN PCIe link into its slice of one array in POSIX shared memory,
# which every rank has mapped and page-locked (cudaHostRegister) piecewise.  Each rank
# first-touches its slice while bound to the CPUs of its GPU's NUMA node, so the pages sit
# behind the root complex the copy arrives at.  No collective on the data path.

def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of a CUDA device's PCIe root (sysfs), or None when unknown."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def _node_cpus(node: int) -> List[int]:
    cpus = []
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            for part in f.read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus += list(range(int(a), int(b) + 1))
                elif part:
                    cpus.append(int(part))
    except (OSError, ValueError):
        pass
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Restricts the calling process to the CPUs of the device's NUMA node (first-touch and
    cudaHostAlloc then place pages there).  Returns the node, or None if nothing was done."""
    import os
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    allowed = set(os.sched_getaffinity(0))
    cpus = [c for c in _node_cpus(node) if c in allowed]
    if not cpus:
        return None
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return node


class SharedHostTable:
    """One uint16[n_off, 4**bits] table in shared host memory, zero-filled, mapped by every
    rank of the group; rank r's row blocks (``triangle_row_blocks``) of every offset are
    page-locked in rank r's process."""

    def __init__(self, mode, group=None, register: bool = True):
        from multiprocessing import shared_memory
        m = _mode_id(mode)
        self.mode = m
        self.bits, self.n_off = MASKED_BITS[m], NUM_OFFSETS[m]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        n = 1 << self.bits
        nbytes = self.n_off * n * n * 2
        name = [None]
        if self.rank == 0:
            self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
            name[0] = self._shm.name
        if self.world > 1:
            dist.broadcast_object_list(name, src=0, group=group)
        if self.rank != 0:
            self._shm = shared_memory.SharedMemory(name=name[0])
            # the creator unlinks; attaching processes must not let the resource tracker
            # unlink the segment a second time at exit
            try:
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self._shm._name, "shared_memory")
            except Exception:   # noqa: BLE001
                pass
        self.array = np.ndarray((self.n_off, n * n), dtype=np.uint16, buffer=self._shm.buf)
        self.base_ptr = self.array.ctypes.data
        self.blocks = triangle_row_blocks(n, self.world, self.rank)
        self._registered = []
        cube = self.array.reshape(self.n_off, n, n)
        self.mine = [cube[o, b:e] for o in range(self.n_off) for b, e in self.blocks]
        for run in self.mine:       # contiguous runs of rows
            run[...] = 0            # first touch: pages land on this process's NUMA node
        self.pinned = False
        if register and torch.cuda.is_available():
            rt = torch.cuda.cudart()
            for run in self.mine:
                err = rt.cudaHostRegister(run.ctypes.data, run.nbytes, 0)
                if int(err) != 0:
                    raise RuntimeError("cudaHostRegister failed: %s" % (err,))
                self._registered.append(run.ctypes.data)
            self.pinned = True
        if self.world > 1:
            dist.barrier(group=group)

    def close(self):
        if self._registered:
            rt = torch.cuda.cudart()
            for p in self._registered:
                rt.cudaHostUnregister(p)
            self._registered = []
        self.mine = None
        self.array = None
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)
        try:
            self._shm.close()
            if self.rank == 0:
                self._shm.unlink()
        except (OSError, BufferError):
            pass


def compute_edit_distance_sharded(mode, lut, host: SharedHostTable,
                                  device_table: Optional[torch.Tensor] = None,
                                  generate_fn: Optional[Callable] = None,
                                  download_fn: Optional[Callable] = None,
                                  barrier: bool = True) -> np.ndarray:
    """compute_edit_distance (make_data_tables.py:111-174) over all ranks of the group, in
    the reference's lower-triangular layout: every rank generates its row blocks on its GPU
    and copies what can be nonzero of them (j < i) into its page-locked slice of ``host``
    over its own PCIe link.  After the closing barrier ``host.array`` holds the whole table
    in every process (rank 0 is the caller that gets "the" array).  No collective moves
    table data.

    ``generate_fn(mode, lut, layout, row_begin, row_end, out)`` and ``download_fn(mode,
    table, host_ptr, row_begin, row_end)`` default to the CUDA generator and
    ``ops.table_download``."""
    m = _mode_id(mode)
    if m != host.mode:
        raise ValueError("host table was made for another mode")
    if generate_fn is None or download_fn is None:
        from . import ops
        generate_fn = generate_fn or ops.table_generate_into
        download_fn = download_fn or ops.table_download
    n = 1 << host.bits
    if device_table is None:
        device_table = torch.empty((host.n_off, n * n), dtype=torch.uint16,
                                   device="cuda" if torch.cuda.is_available() else "cpu")
    for b, e in host.blocks:
        generate_fn(m, lut, _lib.LAYOUT_TRIANGULAR, b, e, device_table)
        download_fn(m, device_table, host.base_ptr, b, e)
    if device_table.is_cuda:
        torch.cuda.current_stream().synchronize()
    if barrier and host.world > 1:
        dist.barrier(group=host.group)
    return host.array
