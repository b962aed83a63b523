#!/usr/bin/env python
"""BASELINE.json configs[4]: grid interpolation forward / backward / double-backward throughput sweep
(2^16..2^24 points x 2-4 levels x 4-16 channels) on one B200, as points/s and as a fraction of the measured
HBM roofline (algorithmic bytes of SURVEY.md section 8d).  ATen's F.grid_sample (NCDHW, the reference's
first-order path) is timed beside it as the kernel-for-kernel bar; the reference's double-backward
extension needs /root/reference + a JIT build and is not available on the GPU box.

    python benchmarks/interp_sweep.py > profiles/rNN_interp_sweep.csv
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from miso_b200 import cuda_gridsample as cu  # noqa: E402
from miso_b200 import field, synth  # noqa: E402

PEAK = 6535.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def levels_for(L, C, dev):
    bound = synth.SCANNET_SUBMAP_BOUND
    scale = {2: 5, 3: 3, 4: 2}[L]
    base = 0.5
    feats = []
    g = torch.Generator(device="cpu").manual_seed(0)
    for l in range(L):
        cell = base / scale ** l
        X, Y, Z = [int(round((b[1] - b[0]) / cell)) for b in bound]
        f = (torch.randn(1, C, Z, Y, X, generator=g) * 1e-2).to(dev).contiguous(memory_format=torch.channels_last_3d)
        feats.append(f)
    return feats, bound


def main():
    dev = torch.device("cuda")
    print("kind,levels,channels,points,distribution,ms,points_per_s,algorithmic_GBps,frac_of_hbm_peak,aten_ms,speedup_vs_aten")
    for (L, C) in [(2, 4), (3, 4), (4, 4), (2, 8), (2, 16)]:
        feats, bound = levels_for(L, C, dev)
        bl = field.bound_to_list(bound)
        b = torch.tensor(bound, device=dev)
        planar = [f.contiguous() for f in feats]   # NCDHW copies for ATen
        for logn in (16, 18, 20, 22, 24):
            N = 1 << logn
            if N * L * C * 4 * 3 > 8e9:
                continue
            for dist in ("uniform", "rays"):
                if dist == "uniform":
                    x = (torch.rand(N, 3, device=dev) * (b[:, 1] - b[:, 0]) + b[:, 0])
                else:
                    mi, _, (R, t) = synth.rgbd_batch(min(N, 1 << 20), num_kf=49, seed=1)
                    ids = mi["sample_frame_ids"][0, :, 0]
                    xw = torch.einsum("nij,nj->ni", R[ids], mi["coords_frame"][0]) + t[ids, :, 0]
                    x = xw.to(dev).repeat((N + xw.shape[0] - 1) // xw.shape[0], 1)[:N].contiguous()
                xn = (2 * (x - b[:, 0]) / (b[:, 1] - b[:, 0]) - 1).reshape(1, N, 1, 1, 3).contiguous()

                # ---- fused multi-level features (one launch) vs ATen per level + cat
                t_f = timeit(lambda: field.field_features_raw(feats, bl, x))
                t_a = timeit(lambda: torch.cat([F.grid_sample(p, xn, align_corners=False, padding_mode="zeros")[0, :, :, 0, 0].T
                                                for p in planar], 1))
                byts = 12 + L * 8 * C * 4 + L * C * 4
                row("fwd_fused_levels", L, C, N, dist, t_f, byts, t_a)

                # ---- per-level plugin on the finest level: fwd, bwd(grid), bwd(grid+coords), double-bwd
                fl = feats[-1].clone().requires_grad_(True)
                pl = planar[-1].clone().requires_grad_(True)
                go = torch.randn(N, C, device=dev)
                g2 = torch.randn(1, N, 1, 1, 3, device=dev)

                def plug_fwd():
                    return cu.grid_sample_3d(fl, xn, padding_mode="zeros", align_corners=False)

                def aten_fwd():
                    return F.grid_sample(pl, xn, align_corners=False, padding_mode="zeros")

                t1 = timeit(lambda: plug_fwd())
                ta1 = timeit(lambda: aten_fwd())
                row("fwd_level", 1, C, N, dist, t1, 12 + 8 * C * 4 + C * 4, ta1)

                gov = go.T.reshape(1, C, N, 1, 1)

                def plug_bwd():
                    fl.grad = None
                    plug_fwd().backward(gov)

                def aten_bwd():
                    pl.grad = None
                    aten_fwd().backward(gov)

                t2 = timeit(plug_bwd) - t1
                ta2 = timeit(aten_bwd) - ta1
                # the caller-side zero-fill of the dense gradient (grid-sized) is part of both timings
                row("bwd_grid_level(+zerofill)", 1, C, N, dist, t2, 12 + C * 4 + 8 * C * 4, ta2)

                xg = xn.clone().requires_grad_(True)

                def plug_dbl():
                    fl.grad = None
                    out = cu.grid_sample_3d(fl, xg, padding_mode="zeros", align_corners=False)
                    (gx,) = torch.autograd.grad(out, xg, gov, create_graph=True)
                    (gx * g2).sum().backward()

                def plug_first():
                    out = cu.grid_sample_3d(fl, xg, padding_mode="zeros", align_corners=False)
                    torch.autograd.grad(out, xg, gov, create_graph=False)

                t3 = timeit(plug_dbl) - timeit(plug_first)
                row("double_bwd_level(+zerofill+glue)", 1, C, N, dist, max(t3, 1e-9), 24 + 2 * C * 4 + 2 * 8 * C * 4 + 12, float("nan"))
                del fl, pl
        del feats, planar
        torch.cuda.empty_cache()


def row(kind, L, C, N, dist, t, bytes_pp, t_aten):
    gbs = bytes_pp * N / t / 1e9
    sp = t_aten / t if t_aten == t_aten else float("nan")
    print(f"{kind},{L},{C},{N},{dist},{t * 1e3:.4f},{N / t:.4e},{gbs:.1f},{gbs / PEAK:.3f},{t_aten * 1e3:.4f},{sp:.2f}", flush=True)


if __name__ == "__main__":
    main()
