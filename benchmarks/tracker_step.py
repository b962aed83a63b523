"""Keyframe tracking step: one LM iteration = normal equations + 6x6 solve + pose update
on 2^14 samples (ncd_quad.yaml:30) of the ScanNet-submap grid.  One-launch kernel vs the previous torch formulation."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from miso_b200 import synth  # noqa: E402
from miso_b200.tracker import Tracker  # noqa: E402


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    poses = synth.keyframe_poses(bench.NUM_KF, synth.SCANNET_SUBMAP_BOUND, seed=55)
    net = bench.build_model(dev, poses, 0)
    mi, gt, (R, t) = synth.rgbd_batch(1 << 14, num_kf=1, seed=3, poses=(poses[0][:1], poses[1][:1]))
    x = mi["coords_frame"][0].to(dev)
    g = gt["sdf"][0].to(dev)
    Rw, tw = R[0].to(dev), t[0].to(dev)
    tr = Tracker(net, loss_type="GM", gm_scale_sdf=0.1, lm_lambda=1e-4)
    out = {"samples": 1 << 14}
    out["normal_equations_kernel_ms"] = timed(lambda: tr.normal_equations(x, g, Rw, tw))
    out["normal_equations_torch_ms"] = timed(lambda: tr.normal_equations_torch(x, g, Rw, tw))
    out["normal_equations_plus_solve_ms"] = timed(lambda: torch.linalg.solve(*[(H, -b) for H, b, _ in [tr.normal_equations(x, g, Rw, tw)]][0]))
    out["lm_iterations_per_s"] = 1e3 / out["normal_equations_plus_solve_ms"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
