#!/usr/bin/env python
"""Strong scaling of the slab-sharded single-grid fit alone (bench.py's `extra.ncd` without the point-chunk arm):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
        benchmarks/slab_scaling.py [p2p|nccl]
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    halo = sys.argv[1] if len(sys.argv) > 1 else "auto"
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = bench.bench_ncd(dev, steps=60, warmup=5, world=world, rank=rank, modes=("slab",), halo=halo)
    if rank == 0:
        keep = {"n_gpus": world, "one_gpu_ms": out["ms_per_step"], "slab_sharded": out.get("slab_sharded")}
        print(json.dumps(keep), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
