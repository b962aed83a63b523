"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(`/root/reference`, CPU, via oracle/ref_loader.py) on small seeded inputs.  Run in the build
container:  python -m oracle.gen_golden
The fixtures store inputs AND reference outputs, so the oracle (tests/test_oracle_golden.py) and the
CUDA path (tests/test_gpu_golden.py) can both be checked where /root/reference does not exist.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from miso_b200 import synth  # noqa: E402  (synthetic input generators only; no kernels involved)
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BOUND = [[-2.0, 2.0], [-1.0, 1.0], [-2.0, 2.0]]
ABOUND = [[-4.0, 4.0], [-2.0, 2.0], [-4.0, 4.0]]


def _np(t):
    return t.detach().cpu().numpy()


def make_ref_gridnet(bound, base_cell=0.5, scale=5, std=0.1, seed=0, num_poses=4, feats=None):
    from grid_opt.models.grid_net import GridNet
    cfg = ref_loader.reference_model_cfg(bound, base_cell_size=base_cell, per_level_scale=scale, num_poses=num_poses)
    net = GridNet(cfg, device="cpu")
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for l, lvl in enumerate(net.features):
            f = feats[l] if feats is not None else torch.randn(lvl.feature.shape, generator=g) * std
            lvl.feature.copy_(f)
    net.decoder.load_state_dict(synth.decoder_weights(8, seed=seed))
    return net


def gen_gridnet():
    import grid_opt.diff as rdiff
    net = make_ref_gridnet(BOUND)
    net.unlock_feature()
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(257, 3, generator=g) * 2 - 1) * torch.tensor([2.3, 1.15, 2.3])
    w = torch.randn(257, 1, generator=g)
    xg = x.clone().requires_grad_(True)
    feat = net.query_feature(xg)
    sdf = net(xg)
    (sdf * w).sum().backward()
    out = {"bound": np.asarray(BOUND, np.float32), "x": _np(x), "w": _np(w), "features": _np(feat), "sdf": _np(sdf),
           "grad_x": _np(xg.grad)}
    for l in range(2):
        out[f"feat{l}"] = _np(net.features[l].feature)
        out[f"grad_feat{l}"] = _np(net.features[l].feature.grad)
    for k, v in net.decoder.state_dict().items():
        out["dec." + k] = _np(v)
    out["gradient3d_autograd"] = _np(rdiff.gradient3d(x.clone().requires_grad_(True), net, "autograd", create_graph=False))
    out["gradient3d_fd"] = _np(rdiff.gradient3d(x.clone(), net, "finitediff", finite_diff_eps=0.024))
    np.savez_compressed(os.path.join(OUT, "gridnet.npz"), **out)


def gen_mapping():
    import grid_opt.loss as rloss
    out = {"bound": np.asarray(BOUND, np.float32)}
    mi, gt, (R, t) = synth.rgbd_batch(900, num_kf=3, bound=BOUND, seed=5, wall_margin=0.3)
    for k, v in {**mi, **gt}.items():
        out["in." + k] = _np(v)
    out["R"], out["t"] = _np(R), _np(t)
    for tag, kw in {"L1fs": dict(loss_type="L1", weight_fs=0.1), "L2fs": dict(loss_type="L2", weight_fs=0.5)}.items():
        net = make_ref_gridnet(BOUND, num_poses=3)
        for k in range(3):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        L = rloss.MisoLossMapping(weight_sdf=1.0, weight_eik=0.0, trunc_dist=0.15, finite_diff_eps=0.024,
                                  grad_method="finitediff", eik_trunc_dist=0.024, **kw)
        ld = L.compute(net, mi, gt)
        total = 0.0
        for v in ld.values():
            total = total + v.mean()
        total.backward()
        for k, v in ld.items():
            out[f"{tag}.{k}"] = _np(v)
        for l in range(2):
            out[f"{tag}.grad_feat{l}"] = _np(net.features[l].feature.grad)
    # eikonal through miso_loss_eikonal directly (MisoLossMappingBase.compute trips over `use_clip`, loss.py:788)
    net = make_ref_gridnet(BOUND, num_poses=3)
    net.unlock_feature()
    g = torch.Generator().manual_seed(12)
    xw = (torch.rand(400, 3, generator=g) * 2 - 1) * torch.tensor([1.9, 0.95, 1.9])
    gts = torch.randn(400, 1, generator=g) * 0.05
    e = rloss.miso_loss_eikonal(net, xw, gts, 0.05, "finitediff", 0.024)
    e.backward()
    out["eik.x"], out["eik.gt"], out["eik.fd_value"] = _np(xw), _np(gts), _np(e)
    for l in range(2):
        out[f"eik.fd_grad_feat{l}"] = _np(net.features[l].feature.grad)
    for l in range(2):
        out[f"feat{l}"] = _np(net.features[l].feature)
    np.savez_compressed(os.path.join(OUT, "mapping.npz"), **out)


def gen_align():
    from grid_opt.models.grid_atlas import GridAtlas
    import grid_opt.align.miso as amiso
    from oracle import oracle as O
    cfg = ref_loader.reference_model_cfg(ABOUND, base_cell_size=1.0, per_level_scale=2, num_poses=1)
    Rt, tt = synth.submap_layout(2, spacing=(4.0, 3.0))
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=4.0, trans_m=0.3)
    atlas = GridAtlas(cfg, device="cpu")
    shapes = O.level_shapes(ABOUND, 1.0, 2, 2, 4)
    out = {"bound": np.asarray(ABOUND, np.float32)}
    for i in range(2):
        atlas.add_submap(torch.tensor(ABOUND), Rp[i], tp[i])
        feats = synth.fill_submap_from_field(shapes, ABOUND, Rt[i], tt[i])
        feats[0][:, :, :, :, :1] = 0
        feats[1][:, :, :, :, :2] = 0
        sm = atlas.get_submap(i)
        with torch.no_grad():
            for l in range(2):
                sm.features[l].feature.copy_(feats[l])
                out[f"sm{i}.feat{l}"] = _np(feats[l])
        out[f"sm{i}.R"], out[f"sm{i}.t"] = _np(Rp[i]), _np(tp[i])
    atlas.precompute_coordinates_for_alignment()
    for l in range(2):
        for i in range(2):
            out[f"coords.sm{i}.level{l}"] = _np(atlas.coordinates_for_alignment(i, l))
    out["intersect01"] = np.asarray(bool(atlas.check_submap_intersection(0, 1)))
    for level in range(2):
        for p in list(atlas.rotation_corrections) + list(atlas.translation_corrections):
            p.grad = None
        ld = amiso.pairwise_loss_latent(atlas, None, 0, 1, level=level, device="cpu")
        (key, val), = ld.items()
        val.backward()
        out[f"L{level}.loss"] = _np(val)
        out[f"L{level}.key"] = np.asarray(key)
        for i in range(2):
            out[f"L{level}.grad_rot{i}"] = _np(atlas.rotation_corrections[i].grad)
            out[f"L{level}.grad_tra{i}"] = _np(atlas.translation_corrections[i].grad)
    np.savez_compressed(os.path.join(OUT, "align.npz"), **out)


def gen_align_variants():
    """pairwise_loss_latent (miso.py:116-211) variants on the atlas of gen_align: align_loss 'L1' and 'cos' at both
    levels, and truncation pruning (trunc_factor, :176-183, needs the source decoder) with the L2 loss."""
    from grid_opt.models.grid_atlas import GridAtlas
    import grid_opt.align.miso as amiso
    from oracle import oracle as O
    cfg = ref_loader.reference_model_cfg(ABOUND, base_cell_size=1.0, per_level_scale=2, num_poses=1)
    Rt, tt = synth.submap_layout(2, spacing=(4.0, 3.0))
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=4.0, trans_m=0.3)
    atlas = GridAtlas(cfg, device="cpu")
    shapes = O.level_shapes(ABOUND, 1.0, 2, 2, 4)
    out = {"bound": np.asarray(ABOUND, np.float32)}
    dec_sd = synth.decoder_weights(8, seed=3)
    for i in range(2):
        atlas.add_submap(torch.tensor(ABOUND), Rp[i], tp[i])
        feats = synth.fill_submap_from_field(shapes, ABOUND, Rt[i], tt[i])
        feats[0][:, :, :, :, :1] = 0
        feats[1][:, :, :, :, :2] = 0
        sm = atlas.get_submap(i)
        sm.decoder.load_state_dict(dec_sd)
        with torch.no_grad():
            for l in range(2):
                sm.features[l].feature.copy_(feats[l])
                out[f"sm{i}.feat{l}"] = _np(feats[l])
        out[f"sm{i}.R"], out[f"sm{i}.t"] = _np(Rp[i]), _np(tp[i])
    for k, v in dec_sd.items():
        out[f"dec.{k}"] = _np(v)
    atlas.precompute_coordinates_for_alignment()

    def run(tag, **kw):
        for p in list(atlas.rotation_corrections) + list(atlas.translation_corrections):
            p.grad = None
        (key, val), = amiso.pairwise_loss_latent(atlas, None, 0, 1, device="cpu", **kw).items()
        val.backward()
        out[f"{tag}.loss"] = _np(val)
        for i in range(2):
            out[f"{tag}.grad_rot{i}"] = _np(atlas.rotation_corrections[i].grad)
            out[f"{tag}.grad_tra{i}"] = _np(atlas.translation_corrections[i].grad)

    for level in range(2):
        for loss in ("L1", "cos"):
            run(f"{loss}.L{level}", level=level, align_loss=loss)
    sm0 = atlas.get_submap(0)
    with torch.no_grad():
        sd = torch.abs(sm0(atlas.coordinates_for_alignment(0, 1)))
    tf = float(torch.median(sd) / sm0.cell_sizes[1])      # keeps about half of the samples
    out["trunc.factor"] = np.asarray(tf, np.float32)
    out["trunc.kept"] = np.asarray(int((sd < tf * sm0.cell_sizes[1]).sum()))
    run("trunc.L1level", level=1, align_loss="L2", trunc_factor=tf)
    np.savez_compressed(os.path.join(OUT, "align_variants.npz"), **out)


def gen_align_sdf():
    """pairwise_loss_sdf (miso.py:14-113) on a 2-submap atlas with two keyframes per submap; L2 / L1 / GM."""
    from grid_opt.models.grid_atlas import GridAtlas
    import grid_opt.align.miso as amiso
    from oracle import oracle as O
    cfg = ref_loader.reference_model_cfg(ABOUND, base_cell_size=1.0, per_level_scale=2, num_poses=2)
    Rt, tt = synth.submap_layout(2, spacing=(4.0, 3.0))
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=4.0, trans_m=0.3)
    atlas = GridAtlas(cfg, device="cpu")
    shapes = O.level_shapes(ABOUND, 1.0, 2, 2, 4)
    Rk, tk = synth.keyframe_poses(4, ABOUND, seed=3, margin=1.0)
    out = {"bound": np.asarray(ABOUND, np.float32), "kf.R": _np(Rk), "kf.t": _np(tk)}
    dec_sd = synth.decoder_weights(8, seed=0)
    for i in range(2):
        atlas.add_submap(torch.tensor(ABOUND), Rp[i], tp[i], num_poses=2)
        for k in range(2):
            atlas.add_kf(Rk[2 * i + k], tk[2 * i + k])
        feats = synth.fill_submap_from_field(shapes, ABOUND, Rt[i], tt[i])
        sm = atlas.get_submap(i)
        sm.decoder.load_state_dict(dec_sd)
        with torch.no_grad():
            for l in range(2):
                sm.features[l].feature.copy_(feats[l] * 0.3)
                out[f"sm{i}.feat{l}"] = _np(feats[l] * 0.3)
        out[f"sm{i}.R"], out[f"sm{i}.t"] = _np(Rp[i]), _np(tp[i])
    for k, v in dec_sd.items():
        out["dec." + k] = _np(v)
    mi, gt, _ = synth.rgbd_batch(1500, num_kf=4, bound=ABOUND, seed=9, poses=(Rk, tk), wall_margin=0.5)
    for k, v in {**mi, **gt}.items():
        out["in." + k] = _np(v)
    loader = [(mi, gt)]
    for loss in ("L2", "L1", "GM"):
        for p in list(atlas.rotation_corrections) + list(atlas.translation_corrections):
            p.grad = None
        ld = amiso.pairwise_loss_sdf(atlas, loader, 0, 1, align_loss=loss, device="cpu")
        (key, val), = ld.items()
        val.backward()
        out[f"{loss}.loss"] = _np(val)
        out[f"{loss}.key"] = np.asarray(key)
        for i in range(2):
            out[f"{loss}.grad_rot{i}"] = _np(atlas.rotation_corrections[i].grad)
            out[f"{loss}.grad_tra{i}"] = _np(atlas.translation_corrections[i].grad)
    # GridAtlas.query_feature / forward (grid_atlas.py:374-399) at world points around both submaps
    for p in list(atlas.rotation_corrections) + list(atlas.translation_corrections):
        p.grad = None
    g = torch.Generator().manual_seed(17)
    xw = (torch.rand(3000, 3, generator=g) * 2 - 1) * torch.tensor([7.0, 3.0, 6.0]) + torch.tensor([2.0, 0.0, 1.0])
    with torch.no_grad():
        out["atlas.xw"], out["atlas.feat"], out["atlas.sdf"] = _np(xw), _np(atlas.query_feature(xw)), _np(atlas(xw))
    np.savez_compressed(os.path.join(OUT, "align_sdf.npz"), **out)


def gen_variants():
    """Loss variants of the other trainers + dense queries: TsdfLoss3D (loss.py:71-144, finite-difference eikonal on
    uniform points), iSDF sdf_loss/tot_loss (loss_isdf.py:280-365), extract_fields (utils_sdf.py:69-86)."""
    import grid_opt.loss as rloss
    import grid_opt.loss_isdf as risdf
    import grid_opt.diff as rdiff
    import packaging.version  # noqa: F401  (utils_sdf.py:63 uses packaging.version without importing the submodule)
    import grid_opt.utils.utils_sdf as rsdf
    out = {"bound": np.asarray(BOUND, np.float32)}
    net = make_ref_gridnet(BOUND)
    net.unlock_feature()
    for l in range(2):
        out[f"feat{l}"] = _np(net.features[l].feature)
    for k, v in net.decoder.state_dict().items():
        out["dec." + k] = _np(v)
    g = torch.Generator().manual_seed(21)
    N = 300
    coords = (torch.rand(N, 3, generator=g) * 2 - 1) * torch.tensor([1.9, 0.95, 1.9])
    gt_sdf = torch.randn(N, 1, generator=g) * 0.1
    valid = (torch.rand(N, 1, generator=g) > 0.3)
    sign = torch.where(gt_sdf > 0.05, torch.ones_like(gt_sdf), torch.where(gt_sdf < -0.05, -torch.ones_like(gt_sdf),
                                                                         torch.zeros_like(gt_sdf)))
    out["tsdf.in_coords"], out["tsdf.in_sdf"], out["tsdf.in_valid"], out["tsdf.in_sign"] = _np(coords), _np(gt_sdf), _np(valid), _np(sign)
    # pick a numpy seed whose eikonal points keep every ReLU pre-activation of every finite-difference evaluation
    # away from 0: at a kink (|h| ~ 1 ulp) the mask, hence the gradient, is decided by rounding noise and no two
    # implementations agree (seed 123 has one h2 = 7.5e-9)
    lin = [m for m in net.decoder.modules() if isinstance(m, torch.nn.Linear)]
    b = np.asarray(BOUND)
    for np_seed in range(124, 400):
        np.random.seed(np_seed)
        pts = torch.from_numpy(np.stack([np.random.uniform(b[d, 0], b[d, 1], N) for d in range(3)], 1)).float()
        offs = torch.cat([torch.zeros(1, 3), 0.024 * torch.eye(3), -0.024 * torch.eye(3)], 0)
        allp = torch.cat([pts + o for o in offs] + [coords], 0)
        with torch.no_grad():
            h1 = lin[0](net.query_feature(allp))
            h2 = lin[1](torch.relu(h1))
        if min(float(h1.abs().min()), float(h2.abs().min())) > 1e-6:
            break
    out["tsdf.np_seed"] = np.asarray(np_seed)
    np.random.seed(np_seed)
    L = rloss.TsdfLoss3D(grad_method="finitediff", finite_diff_eps=0.024)
    ld = L.compute(net, {"coords": coords[None]}, {"sdf": gt_sdf[None], "sdf_valid": valid[None], "sdf_sign": sign[None]})
    sum(ld.values()).backward()
    for k, v in ld.items():
        out[f"tsdf.{k}"] = _np(v)
    for l in range(2):
        out[f"tsdf.grad_feat{l}"] = _np(net.features[l].feature.grad)
        net.features[l].feature.grad = None
    # iSDF: shapes as in compute_default (pc (1,N,3), bounds (1,N,1), sdf (N,1), grad (1,N,3))
    bounds = torch.rand(N, 1, generator=g) * 0.4
    pc = coords.clone()
    sdf = net(pc)
    gradv = rdiff.gradient3d(pc, net, "finitediff", finite_diff_eps=0.024)
    mat, fs = risdf.sdf_loss(sdf, bounds[None], 0.15, loss_type="L1")
    eik = torch.abs(gradv[None].norm(2, dim=-1) - 1)
    total, _, _ = risdf.tot_loss(mat, None, eik, fs, bounds[None], 0.1, 5.38, 0.0, 0.268)
    total.backward()
    out["isdf.bounds"], out["isdf.total"] = _np(bounds), _np(total)
    for l in range(2):
        out[f"isdf.grad_feat{l}"] = _np(net.features[l].feature.grad)
    bt = torch.tensor(BOUND)
    out["fields.u"] = rsdf.extract_fields(bt[:, 0], bt[:, 1], 20, lambda p: net(p))
    np.savez_compressed(os.path.join(OUT, "variants.npz"), **out)


def gen_tracker():
    """Tracker.lm_step (grid_opt/slam/tracker.py:148-212) executed verbatim on CPU with a minimal stand-in for `self`
    (the method only touches the attributes listed here): pose corrections after ONE LM step, L2 and GM, with and
    without the |gt| < trunc filter."""
    import types
    import grid_opt.slam.tracker as rtr
    out = {"bound": np.asarray(BOUND, np.float32)}
    g = torch.Generator().manual_seed(31)
    N = 1500
    xf = (torch.rand(N, 3, generator=g) - 0.5) * torch.tensor([2.0, 1.0, 2.0])
    gt_sdf = torch.randn(N, 1, generator=g) * 0.05
    from oracle import oracle as O
    Rwf = O.so3_exp_map(torch.tensor([[0.1, -0.2, 0.05]]))[0]
    twf = torch.tensor([[0.1], [0.05], [-0.1]])
    out["coords_frame"], out["gt_sdf"], out["Rwf"], out["twf"] = _np(xf), _np(gt_sdf), _np(Rwf), _np(twf)
    mi = {"coords_frame": xf[None], "sample_frame_ids": torch.zeros(1, N, 1, dtype=torch.long), "weights": torch.ones(1, N, 1)}
    gt = {"sdf": gt_sdf[None], "sdf_valid": torch.ones(1, N, 1, dtype=torch.bool), "sdf_signs": torch.zeros(1, N, 1)}
    first = True
    for loss_type in ("L2", "GM"):
        for trunc in (None, 0.04):
            net = make_ref_gridnet(BOUND, num_poses=1)
            net.set_initial_kf_pose(0, Rwf, twf, kf_key="KF0")
            if first:
                for l in range(2):
                    out[f"feat{l}"] = _np(net.features[l].feature)
                for k, v in net.decoder.state_dict().items():
                    out["dec." + k] = _np(v)
                first = False
            fake = types.SimpleNamespace(dataset=types.SimpleNamespace(select_keyframes=lambda kfs: None),
                                         train_loader=[(mi, gt)], cfg={"device": "cpu"}, trunc_dist=trunc, grid=net,
                                         loss_type=loss_type, gm_scale_sdf=0.1, lm_lambda=1e-4)
            fake.residual_weights = types.MethodType(rtr.Tracker.residual_weights, fake)
            info = rtr.Tracker.lm_step(fake, 0)
            tag = f"{loss_type}.{'trunc' if trunc else 'all'}"
            out[f"{tag}.delta_R"] = _np(net.rotation_corrections[0])
            out[f"{tag}.delta_t"] = _np(net.translation_corrections[0])
            out[f"{tag}.info"] = np.asarray([info["delta_R_deg"], info["delta_t_norm"], info["grad_norm"], info["fov_overlap"]])
    np.savez_compressed(os.path.join(OUT, "tracker.npz"), **out)


def gen_align_loop():
    """The whole alignment loop run by the reference: align_multiple_submaps_hierarchical (miso.py:217-322) ->
    generic_align_multiple_submaps (base.py:89-163) with pairwise_loss_latent at levels 0 and 1, 5 iterations each
    (`while iter <= num_iters`), Adam lr 1e-2, intersection test on, on a 3-submap atlas: final pose corrections."""
    from grid_opt.models.grid_atlas import GridAtlas
    import grid_opt.align.miso as amiso
    from oracle import oracle as O
    cfg = ref_loader.reference_model_cfg(ABOUND, base_cell_size=1.0, per_level_scale=2, num_poses=1)
    Rt, tt = synth.submap_layout(3, spacing=(4.0, 3.0))
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=4.0, trans_m=0.3)
    atlas = GridAtlas(cfg, device="cpu")
    shapes = O.level_shapes(ABOUND, 1.0, 2, 2, 4)
    out = {"bound": np.asarray(ABOUND, np.float32)}
    for i in range(3):
        atlas.add_submap(torch.tensor(ABOUND), Rp[i], tp[i])
        feats = synth.fill_submap_from_field(shapes, ABOUND, Rt[i], tt[i])
        with torch.no_grad():
            for l in range(2):
                atlas.get_submap(i).features[l].feature.copy_(feats[l])
                out[f"sm{i}.feat{l}"] = _np(feats[l])
        out[f"sm{i}.R"], out[f"sm{i}.t"] = _np(Rp[i]), _np(tp[i])
    # the loop's PerfTimer records CUDA events (utils.py) -- not part of the algorithm and unusable without a GPU
    import grid_opt.align.base as rbase

    class _CpuTimer:
        def __init__(self, activate=True):
            pass

        def reset(self):
            pass

        def check(self):
            return 0.0, 0.0
    rbase.utils.PerfTimer = _CpuTimer
    amiso.align_multiple_submaps_hierarchical(atlas, [0], level_iters=4, lr=1e-2, latent_levels=[0, 1], skip_finetune=True,
                                              device="cpu", verbose=False)
    for i in range(3):
        out[f"final.rot{i}"] = _np(atlas.rotation_corrections[i])
        out[f"final.tra{i}"] = _np(atlas.translation_corrections[i])
    np.savez_compressed(os.path.join(OUT, "align_loop.npz"), **out)


def main():
    ref_loader.load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    np.random.seed(0)
    gens = {"gridnet": gen_gridnet, "mapping": gen_mapping, "align": gen_align, "align_variants": gen_align_variants,
            "align_sdf": gen_align_sdf, "variants": gen_variants, "tracker": gen_tracker, "align_loop": gen_align_loop}
    for name in (sys.argv[1:] or list(gens)):     # `python oracle/gen_golden.py align_variants` regenerates one fixture
        gens[name]()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
