"""What bounds the mapping step: the two memory halves of the fused kernel timed on their own, on the bench
workload (2^20 RGB-D ray samples, ScanNet-submap grid), through the C-ABI:

  * miso_sdf_forward  (gather of both levels + decoder + Jacobian, no scatter),
  * miso_sdf_backward (pure scatter: 16 x red.global.add.v4.f32 per point, reads xw/jac/a/v),
  * miso_field_features (pure gather).

Usage: python benchmarks/scatter_probe.py [--points N]   -> one JSON line."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from miso_b200 import field as F  # noqa: E402
from miso_b200.loss import MisoLossMapping  # noqa: E402


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=bench.N_POINTS)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    bench.N_POINTS = args.points
    batches, poses = bench.make_host_batches(0)
    net = bench.build_model(dev, poses, 0)
    mi, gt = batches[0]
    x = mi["coords_frame"][0].to(dev)
    ids = mi["sample_frame_ids"][0, :, 0].to(dev)
    loss = MisoLossMapping(**bench.LOSS_CFG)
    frames = loss._frames(net, ids)
    spec = net.fused_spec()
    feats = net.level_tensors()
    grads = [torch.zeros_like(f) for f in feats]
    N = x.shape[0]
    sdf, jac, gradx, xw = F.sdf_forward_raw(feats, spec, x, frames, want_xw=True)
    a = torch.randn(N, device=dev) / N
    v = torch.randn(N, 3, device=dev) / N
    res = {"points": N}
    res["sdf_forward_ms"] = timed(lambda: F.sdf_forward_raw(feats, spec, x, frames, want_xw=True))
    res["sdf_backward_scatter_ms"] = timed(lambda: F.sdf_backward_raw(feats, grads, spec, xw, jac, a, v))
    res["sdf_backward_scatter_a_only_ms"] = timed(lambda: F.sdf_backward_raw(feats, grads, spec, xw, jac, a, None))
    res["field_features_gather_ms"] = timed(lambda: F.field_features_raw(feats, spec.bound, xw))
    res["red128_per_s"] = 16 * N / (res["sdf_backward_scatter_ms"] * 1e-3)
    res["red128_cycles_per_lane_per_sm"] = (res["sdf_backward_scatter_ms"] * 1e-3) * 1.965e9 * 148 / (16 * N)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
