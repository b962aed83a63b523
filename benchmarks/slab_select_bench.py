#!/usr/bin/env python
"""Device time of miso_slab_select on the NCD quad batch (2^22 LiDAR samples) for rank 0 of `world` emulated slabs,
on ONE GPU (the kernel has no communication).  python benchmarks/slab_select_bench.py [world]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from miso_b200 import _lib, sharded_fit as sf, synth  # noqa: E402
from miso_b200.loss import MisoLossMapping  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda", 0)
    mi, gt, poses = synth.lidar_batch(bench.NCD_POINTS, num_kf=bench.NCD_KF, seed=3)
    dmi = {k: v.to(dev) for k, v in mi.items()}
    dgt = {k: v.to(dev) for k, v in gt.items()}
    net = bench.build_ncd_model(dev, poses)
    out = {}
    for rank in range(world):
        fit = sf.SlabShardedFit(net, MisoLossMapping(**bench.NCD_LOSS), lr=1e-3, rank=rank, world=world)
        fit.calibrate(dmi)
        lib = _lib.load()
        L = fit.loss
        coords = dmi["coords_frame"][0].contiguous()
        ids = dmi["sample_frame_ids"][0, :, 0]
        sdf = dgt["sdf"][0].reshape(-1).contiguous()
        valid = dgt["sdf_valid"][0].reshape(-1).contiguous().view(torch.uint8)
        sign = dgt["sdf_signs"][0].reshape(-1).contiguous()
        w = dmi["weights"][0].reshape(-1).contiguous()
        N = coords.shape[0]
        b = fit._buffers(N, dev, True)
        fr = L._frames(net, ids).struct()
        stream = _lib.stream_ptr(dev)

        def run():
            _lib.check(lib.miso_slab_select(
                C.byref(fr), coords.data_ptr(), N, float(fit.zmin), float(fit.zmax), fit.Z, fit.axis, fit.zb, fit.ze,
                sdf.data_ptr(), valid.data_ptr(), sign.data_ptr(), w.data_ptr(), b["x"].data_ptr(), b["ids"].data_ptr(),
                b["sdf"].data_ptr(), b["valid"].data_ptr(), b["sign"].data_ptr(), b["w"].data_ptr(),
                b["count"].data_ptr(), stream), "slab_select")

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        kept = int(b["count"].item())
        bytes_ = N * 20 + kept * (13 + 37)
        out[f"rank{rank}"] = {"axis": fit.axis, "slab": [fit.zb, fit.ze], "kept": kept, "ms": ms, "GBps": bytes_ / ms / 1e6}
    print(json.dumps({"world": world, "N": bench.NCD_POINTS, "select": out}))


if __name__ == "__main__":
    main()
