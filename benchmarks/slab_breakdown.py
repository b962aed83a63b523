#!/usr/bin/env python
"""Per-phase device time of one SlabShardedFit step on the NCD quad workload (torchrun --nproc-per-node N)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from miso_b200 import _lib, dist as mdist, sharded_fit as sf, synth  # noqa: E402
from miso_b200.loss import MisoLossMapping  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mi, gt, poses = synth.lidar_batch(bench.NCD_POINTS, num_kf=bench.NCD_KF, seed=3)
    dmi = {k: v.to(dev) for k, v in mi.items()}
    dgt = {k: v.to(dev) for k, v in gt.items()}
    net = bench.build_ncd_model(dev, poses)
    fit = sf.SlabShardedFit(net, MisoLossMapping(**bench.NCD_LOSS), lr=1e-3)
    fit.calibrate(dmi)
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    fit.mark = mark      # SlabShardedFit reports its phase boundaries
    lib = _lib.load()
    acc = {}
    for it in range(12):
        marks.clear()
        mark("start")
        fit.step(dmi, dgt)
        torch.cuda.synchronize()
        if it >= 4:
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1) / 8
            acc["total"] = acc.get("total", 0.0) + marks[0][1].elapsed_time(marks[-1][1]) / 8
    # back-to-back launches of single phases (no host gaps)
    import ctypes as C
    from miso_b200 import field as _field
    from miso_b200.loss import _flat_f32, _flat_u8
    L = fit.loss
    coords = _field._prep_x(dmi["coords_frame"][0])
    ids = dmi["sample_frame_ids"][0, :, 0]
    sdf, valid, sign = _flat_f32(dgt["sdf"][0]), _flat_u8(dgt["sdf_valid"][0]), _flat_f32(dgt["sdf_signs"][0])
    w = _flat_f32(dmi["weights"][0])
    b = fit._bufs
    fr = L._frames(net, ids).struct()

    def t_loop(fn, n=10):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    acc["select_only"] = t_loop(lambda: lib.miso_slab_select(
        C.byref(fr), coords.data_ptr(), coords.shape[0], float(fit.zmin), float(fit.zmax), fit.Z, fit.axis, fit.zb, fit.ze,
        sdf.data_ptr(), valid.data_ptr(), sign.data_ptr(), w.data_ptr(), b["x"].data_ptr(), b["ids"].data_ptr(),
        b["sdf"].data_ptr(), b["valid"].data_ptr(), b["sign"].data_ptr(), b["w"].data_ptr(), b["count"].data_ptr(),
        _lib.stream_ptr(dev)))
    feats = net.level_tensors()
    n_sl = (fit.ze - fit.zi) * fit.plane_elems
    off = fit.zi * fit.plane_elems * 4
    acc["adam_slab_only_zero_grad"] = t_loop(lambda: lib.miso_adam_step_dev(
        feats[1].data_ptr() + off, feats[1].grad.data_ptr() + off, fit.exp_avg.data_ptr(), fit.exp_avg_sq.data_ptr(),
        fit.touched.data_ptr(), n_sl, fit.lr, 0.9, 0.999, fit.eps, fit.step_dev.data_ptr(), fit.scalars.data_ptr(), None, 1,
        _lib.stream_ptr(dev)))
    acc["touched_fraction_of_slab"] = float(sum(bin(int(x) & 0xffffffff).count("1") for x in fit.touched[:200000].tolist())) / (200000 * 32)
    print(json.dumps({"rank": rank, "world": world, "slab": [fit.zb, fit.ze], "own": int(fit._bufs["count"].item()), "ms": acc}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
