"""Multi-GPU partitioning of the hot path: one process per GPU under
torchrun, `torch.distributed` (NCCL over NVLink/NVSwitch on the box, gloo in the CPU tests).

  * build_submaps / local optimisation: submaps are independent GridNets (demo/build_submaps.py:133-134)
    -> submap i on rank i % world, NO data-path collective; submaps are gathered to rank 0 in id order
    only to assemble the GridAtlas (ModuleList order = submap id, grid_atlas.py:145-150).
  * large single-grid fit: points sharded per rank, loss = mean over ALL points (loss.py:634) -> every rank
    scales by N_local/N_total and the dense grid gradients are summed with one all_reduce per level
    before Adam (every rank then applies the identical update).
  * alignment: total_loss = sum over pairs (base.py:146) -> pairs round-robin (or cost-balanced) over ranks, one all_reduce of
    the per-submap pose gradients (<= 16 x 6 floats) per iteration.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so the pinned host staging buffers it
    allocates afterwards (first touch) and its H2D copies stay on the local socket.  With one process per GPU and 8
    GPUs the default placement sends most of the 8 x 53 GB/s of host reads across the socket interconnect.
    Returns the node, or None when the topology cannot be read (then nothing is changed)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def submaps_for_rank(num_submaps: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Round-robin submap assignment: submap i -> rank i % world."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return [i for i in range(num_submaps) if i % world_size == rank]


def owner_of_submap(submap_id: int, world_size: Optional[int] = None) -> int:
    _, w = world()
    return submap_id % (w if world_size is None else world_size)


def shard_points(n: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous [begin, end) chunk of an n-point batch for this rank (chunks of a Morton-sorted batch
    keep per-GPU locality)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(n, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def pair_filter(rank: Optional[int] = None, world_size: Optional[int] = None) -> Callable[[int, Tuple[int, int]], bool]:
    """Round-robin pair ownership for generic_align_multiple_submaps(pair_filter=...)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return lambda i, pair: i % world_size == rank


def balanced_pair_owner(costs: Sequence[float], world_size: Optional[int] = None) -> List[int]:
    """Cost-balanced pair ownership ("or cost-balanced by M_valid"): longest-processing-time
    greedy -- pairs sorted by decreasing cost (e.g. alignment samples of the pair, 0 when the submaps do not
    intersect) go to the currently lightest rank.  Deterministic (ties by pair index), identical on every rank.
    Returns owner[i] for every pair."""
    _, w = world()
    world_size = w if world_size is None else world_size
    load = [0.0] * world_size
    owner = [0] * len(costs)
    for i in sorted(range(len(costs)), key=lambda k: (-float(costs[k]), k)):
        r = min(range(world_size), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += float(costs[i])
    return owner


def balanced_pair_filter(costs: Sequence[float], rank: Optional[int] = None,
                         world_size: Optional[int] = None) -> Callable[[int, Tuple[int, int]], bool]:
    r, _ = world()
    rank = r if rank is None else rank
    owner = balanced_pair_owner(costs, world_size)
    return lambda i, pair: owner[i] == rank


def allreduce_sum_(tensors: Sequence[torch.Tensor]) -> None:
    """In-place SUM over ranks.  Small tensors are flattened into one buffer -> one collective
    (latency-bound: pose gradients are ~100 floats); large ones (dense grid gradients) go one by one so
    the coarse level's reduction can overlap the fine level's."""
    _, w = world()
    if w == 1 or not tensors:
        return
    small = [t for t in tensors if t.numel() <= 65536]
    large = [t for t in tensors if t.numel() > 65536]
    handles = [dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True) for t in large]
    if small:
        flat = torch.cat([t.reshape(-1) for t in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        off = 0
        for t in small:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
    for h in handles:
        h.wait()


def gather_submaps_to_rank0(local: dict, num_submaps: int):
    """{submap_id: state_dict} per rank -> ordered list on rank 0 (None elsewhere)."""
    r, w = world()
    if w == 1:
        return [local[i] for i in range(num_submaps)]
    gathered = [None] * w
    dist.all_gather_object(gathered, local)
    if r != 0:
        return None
    merged = {}
    for d in gathered:
        merged.update(d)
    return [merged[i] for i in range(num_submaps)]


def sharded_loss_scale(n_local: int, n_total: int) -> float:
    """A rank's mean-over-local-points loss must be weighted by N_local/N_total so the summed gradients
    equal the single-GPU gradient of the mean over all points."""
    return float(n_local) / float(n_total)
