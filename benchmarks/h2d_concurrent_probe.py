"""How much pinned-host -> device bandwidth do k concurrent ranks get on this box?  Explains the e2e numbers of
`bench.py --gpus N`: each rank feeds its GPU 34.6 MB per step from pinned memory (53 GB/s alone).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        benchmarks/h2d_concurrent_probe.py

For k = 1, 2, 4, 8 (<= world) the first k ranks copy a 64 MiB pinned buffer to their GPU in a loop while the others
idle; rank 0 prints one JSON line with per-rank and aggregate GB/s, the NUMA node sysfs reports for every GPU and the
CPU affinity each rank ended up with (miso_b200.dist.bind_to_gpu_numa_node).
"""
import glob
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miso_b200 import dist as mdist  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    bind = os.environ.get("MISO_PROBE_BIND", "1") == "1"
    node = mdist.bind_to_gpu_numa_node(local) if bind else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 64 << 20
    src = torch.empty(n, dtype=torch.uint8).pin_memory()
    src.fill_(1)
    dst = torch.empty(n, dtype=torch.uint8, device="cuda")
    p = torch.cuda.get_device_properties(local)
    sysfs = f"/sys/bus/pci/devices/{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0/numa_node"
    try:
        gpu_node = int(open(sysfs).read())
    except OSError:
        gpu_node = None
    rows = []
    ks = [k for k in (1, 2, 4, 8) if k <= world]
    for k in ks:
        active = rank < k
        for _ in range(2):
            if active:
                dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        if active:
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        gbs = n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9 if active else 0.0
        t = torch.tensor([gbs], device="cuda")
        if world > 1:
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            per = [float(o) for o in out]
        else:
            per = [gbs]
        rows.append({"active_ranks": k, "per_rank_GBs": [round(v, 1) for v in per[:k]], "aggregate_GBs": round(sum(per), 1)})
    info = torch.tensor([gpu_node if gpu_node is not None else -9, node if node is not None else -9,
                         len(os.sched_getaffinity(0))], device="cuda")
    if world > 1:
        infos = [torch.zeros_like(info) for _ in range(world)]
        dist.all_gather(infos, info)
    else:
        infos = [info]
    if rank == 0:
        print(json.dumps({"world": world, "bind_requested": bind, "host_numa_nodes": len(glob.glob("/sys/devices/system/node/node[0-9]*")),
                          "cpus": os.cpu_count(),
                          "gpu_numa_node_sysfs/bound_node/affinity_cpus": [[int(v) for v in i.tolist()] for i in infos],
                          "copies": rows}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
