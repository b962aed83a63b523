"""Dense forward queries: extract_fields (utils_sdf.py:69-86) at 256^3 on the
ScanNet-submap grid -- one fused forward launch per slab vs the reference's 16^3-point chunk loop (same kernels,
4096 points per launch, timed on a 1/64 sample of the chunks).  Prints one JSON line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from miso_b200 import synth  # noqa: E402
from miso_b200.utils_sdf import extract_fields  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    poses = synth.keyframe_poses(bench.NUM_KF, synth.SCANNET_SUBMAP_BOUND, seed=55)
    net = bench.build_model(dev, poses, 0)
    b = torch.tensor(synth.SCANNET_SUBMAP_BOUND)
    res = 256
    extract_fields(b[:, 0], b[:, 1], 64, net, device=dev)          # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    u = extract_fields(b[:, 0], b[:, 1], res, net, device=dev)
    t1 = time.perf_counter()
    # device-only time of the fused forward on 2^23 points: (a) one lattice slab, the access pattern of
    # extract_fields (x-major 'ij' meshgrid, z fastest), (b) uniform random points (no locality: every 16-byte corner
    # costs a 32-byte L2 sector)
    X = torch.linspace(-10.0, 10.0, 128, device=dev)
    Y = torch.linspace(-5.0, 5.0, 256, device=dev)
    Z = torch.linspace(-10.0, 10.0, 256, device=dev)
    xx, yy, zz = torch.meshgrid(X, Y, Z, indexing="ij")
    lattice = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1).contiguous()
    pts = (torch.rand(1 << 23, 3, device=dev) * 2 - 1) * torch.tensor([10.0, 5.0, 10.0], device=dev)

    def time_fwd(p):
        with torch.no_grad():
            net(p)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(5):
                net(p)
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / 5

    fwd_lattice_ms = time_fwd(lattice)
    fwd_ms = time_fwd(pts)
    with torch.no_grad():
        # the reference's chunking: 16^3 points per call + a device->host copy per chunk (sampled)
        chunk = pts[:4096]
        n_chunks = (res // 16) ** 3
        sample = max(1, n_chunks // 64)
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        for _ in range(sample):
            net(chunk).cpu()
        c1 = time.perf_counter()
    out = {"resolution": res, "points": res ** 3, "extract_fields_s": t1 - t0,
           "extract_fields_points_per_s": res ** 3 / (t1 - t0), "fused_forward_lattice_ms_per_2^23": fwd_lattice_ms,
           "fused_forward_lattice_points_per_s": (1 << 23) / (fwd_lattice_ms * 1e-3),
           "forward_lattice_frac_of_hbm_peak": 272 * (1 << 23) / (fwd_lattice_ms * 1e-3) / 1e9 / bench.measured_hbm_peak()[0],
           "fused_forward_ms_per_2^23": fwd_ms,
           "fused_forward_points_per_s": (1 << 23) / (fwd_ms * 1e-3),
           "forward_bytes_per_point": 272, "forward_frac_of_hbm_peak": 272 * (1 << 23) / (fwd_ms * 1e-3) / 1e9 / bench.measured_hbm_peak()[0],
           "chunked_16cubed_s_extrapolated": (c1 - c0) / sample * n_chunks, "finite": bool(torch.isfinite(torch.from_numpy(u)).all())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
