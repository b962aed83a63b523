#!/usr/bin/env python
"""BASELINE.json configs[4]: grid interpolation forward / backward / double-backward throughput sweep
(2^16..2^24 points x 2-4 levels x 4-16 channels) on one B200, as points/s and as a fraction of the measured
HBM roofline (algorithmic bytes of SURVEY.md section 8d).  ATen's F.grid_sample (NCDHW, the reference's
first-order path) is timed beside it as the kernel-for-kernel bar, and -- when oracle/_ref/gridsample_grad2.so was
built (oracle/build_ref.py: the reference's own extension, unmodified, for sm_100a) -- the reference's double-backward
kernel `grid_sampler_3d_grad2_kernel` (third_party/cuda_gridsample_grad2/gridsample_cuda.cu:212-533), both through
its autograd plugin and launch-for-launch against miso_grid_sample3d_bwd_bwd.

    python benchmarks/interp_sweep.py > profiles/rNN_interp_sweep.csv
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from miso_b200 import _lib  # noqa: E402
from miso_b200 import cuda_gridsample as cu  # noqa: E402
from miso_b200 import field, synth  # noqa: E402
from oracle import build_ref, ref_gpu  # noqa: E402  (benchmark baseline only)

REF_EXT = build_ref.load_module()

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
    print("kind,levels,channels,points,distribution,ms,points_per_s,algorithmic_GBps,frac_of_hbm_peak,baseline_ms,speedup_vs_baseline")
    print("# baseline = ATen F.grid_sample for fwd/bwd rows, the reference's grad2 extension for double_bwd rows", flush=True)
    quick = os.environ.get("MISO_SWEEP_QUICK", "0") == "1"
    for (L, C) in ([(2, 4)] if quick else [(2, 4), (3, 4), (2, 8), (2, 16)]):
        feats, bound = levels_for(L, C, dev)
        bl = field.bound_to_list(bound)
        b = torch.tensor(bound, device=dev)
        planar = [f.contiguous() for f in feats]   # NCDHW copies for ATen
        for logn in ((20, 22) if quick else (16, 18, 20, 22, 24)):
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
                dbl_bytes = 24 + 2 * C * 4 + 2 * 8 * C * 4 + 12
                t3_ref = float("nan")
                if REF_EXT is not None:
                    def ref_dbl():
                        pl.grad = None
                        out = ref_gpu.grid_sample_3d(pl, xg, padding_mode="zeros", align_corners=False)
                        (gx,) = torch.autograd.grad(out, xg, gov, create_graph=True)
                        (gx * g2).sum().backward()

                    def ref_first():
                        out = ref_gpu.grid_sample_3d(pl, xg, padding_mode="zeros", align_corners=False)
                        torch.autograd.grad(out, xg, gov, create_graph=False)

                    t3_ref = timeit(ref_dbl) - timeit(ref_first)
                row("double_bwd_level(+zerofill+glue)", 1, C, N, dist, max(t3, 1e-9), dbl_bytes, t3_ref)

                # launch for launch: gg_grid given, outputs gg_output + g_input (accumulated) -- the eikonal pattern
                lib = _lib.load()
                gg_out = torch.empty(1, N, C, device=dev)
                g_in = torch.zeros_like(fl.detach())
                xnc, g2c = xn.reshape(1, N, 3).contiguous(), g2.reshape(1, N, 3).contiguous()
                fd = fl.detach()
                args = (_lib.F32, None, None, g2c.data_ptr(), gov.data_ptr(), _lib.i64(gov.reshape(1, C, N).stride()),
                        fd.data_ptr(), _lib.i64(fd.shape), _lib.i64(fd.stride()), xnc.data_ptr(), N, gg_out.data_ptr(),
                        _lib.i64([gg_out.stride(0), gg_out.stride(2), gg_out.stride(1)]), g_in.data_ptr(),
                        _lib.i64(g_in.stride()), None, 0, 0, _lib.stream_ptr(dev))

                def ours_kernel():
                    _lib.check(lib.miso_grid_sample3d_bwd_bwd(*args), "grid_sample3d_bwd_bwd")

                t4 = timeit(ours_kernel)
                t4_ref = float("nan")
                if REF_EXT is not None:
                    zeros_in = torch.zeros_like(pl.detach())
                    govc = gov.contiguous()
                    pld = pl.detach()

                    def ref_kernel():   # allocates + zero-fills its three outputs inside, as the reference does
                        REF_EXT.grad2_3d(zeros_in, g2, govc, pld, xn, False, False)

                    t4_ref = timeit(ref_kernel)
                    del zeros_in
                row("double_bwd_kernel_only", 1, C, N, dist, t4, dbl_bytes, t4_ref)
                del fl, pl, g_in, gg_out
        del feats, planar
        torch.cuda.empty_cache()


def row(kind, L, C, N, dist, t, bytes_pp, t_aten):
    gbs = bytes_pp * N / t / 1e9
    sp = t_aten / t if t_aten == t_aten else float("nan")
    print(f"{kind},{L},{C},{N},{dist},{t * 1e3:.4f},{N / t:.4e},{gbs:.1f},{gbs / PEAK:.3f},{t_aten * 1e3:.4f},{sp:.2f}", flush=True)


if __name__ == "__main__":
    main()
