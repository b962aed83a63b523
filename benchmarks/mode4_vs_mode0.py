#!/usr/bin/env python
"""Step-kernel time of the device-count variant (mode 4, used by SlabShardedFit after miso_slab_select) against the
plain step (mode 0) on the same NCD quad batch, one GPU; also on the compacted half batch of an emulated 2-rank y-slab."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from miso_b200 import loss as mloss, sharded_fit as sf, synth  # noqa: E402
from miso_b200.loss import MisoLossMapping  # noqa: E402
from miso_b200.trainer import GridTrainer  # noqa: E402


def kernel_ms(step, n=30, warm=5):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    mloss.PROFILE_EVENTS = []
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in mloss.PROFILE_EVENTS]))
    mloss.PROFILE_EVENTS = None
    return ms


def main():
    dev = torch.device("cuda", 0)
    mi, gt, poses = synth.lidar_batch(bench.NCD_POINTS, num_kf=bench.NCD_KF, seed=3)
    dmi = {k: v.to(dev) for k, v in mi.items()}
    dgt = {k: v.to(dev) for k, v in gt.items()}
    out = {}
    net = bench.build_ncd_model(dev, poses)
    tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, net, MisoLossMapping(**bench.NCD_LOSS), None, device=dev)
    out["mode0_full_batch_ms"] = kernel_ms(lambda: tr.train_step(dmi, dgt))
    del tr, net
    net = bench.build_ncd_model(dev, poses)
    fit = sf.SlabShardedFit(net, MisoLossMapping(**bench.NCD_LOSS), lr=1e-3, rank=0, world=1)
    fit.calibrate(dmi)
    out["mode4_full_batch_ms"] = kernel_ms(lambda: fit.step(dmi, dgt))
    del fit, net
    for rank in (0, 1):
        net = bench.build_ncd_model(dev, poses)
        fit = sf.SlabShardedFit(net, MisoLossMapping(**bench.NCD_LOSS), lr=1e-3, rank=rank, world=2)
        fit.calibrate(dmi)
        fit._exchange_and_update = lambda *a, **k: None      # selection + kernel only (no process group here)
        out[f"mode4_half_batch_rank{rank}_ms"] = kernel_ms(lambda: fit.step(dmi, dgt))
        out[f"half_batch_rank{rank}_samples"] = int(fit._bufs["count"].item())
        out["slab_axis"] = fit.axis
        del fit, net
    print(json.dumps(out))


if __name__ == "__main__":
    main()
