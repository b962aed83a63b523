"""GPU parity of the fused mapping step AT BASELINE.json SIZES: the persistent
multi-tile loop of the headline kernel (`mapping_step_tc2_kernel`: 148 CTAs x 4 groups, so 2^18 points are
already 3-4 tiles per group and 2^20 points 14), its one-tile-ahead point staging and TMEM / A_lo reuse are
value-checked against the oracle, not only run for finite losses.

  * configs[0]/[1]: 2^18 and 2^20 RGB-D-sampled points on the ScanNet-submap grid (40x20x40 + 200x100x200),
    L1 sdf + 0.1 free space + 0.5 second-order eikonal (scannet.yaml:43-49 with grad_method autograd);
  * configs[3]: 2^22 LiDAR-sampled points on the Newer-College-quad grid (20x90x90 + 100x450x450, 324 MB fine
    level), L2 sdf + 0.5 free space, trunc 0.5 (ncd_quad.yaml:42-46), plus the eikonal term;
  * a grid with >= 2^31 elements (the 64-bit-offset route, gridsample_cuda.cu:628-660) against the oracle on the
    crop of the grid the points touch (trilinear interpolation is local);
  * every kernel variant (3 tiles in flight, unpaired lanes, one thread per point, SIMT decoder, forced 64-bit
    offsets) on a multi-tile batch.

Tolerances (BASELINE.json north_star): loss terms 1e-5 relative, grid gradients 1e-4 relative.
The oracle runs chunked (tests/helpers.py::oracle_mapping_chunked); every chunk is the reference's op sequence.

Kinks.  d(total)/d(grid) is discontinuous where a decoder ReLU pre-activation, the L1 residual or the free-space
branch difference crosses zero; a sample within rounding distance of a kink gets a different one-sided derivative
from two correct float32 implementations that merely sum in another order, and at 2^18..2^22 samples ONE such flip is
~1/sqrt(N) = 1e-4..1e-3 of the gradient norm (measured: the reference's float32 op sequence itself sits 4e-4..3e-3
from its own float64 evaluation).  So every size is checked twice:
  * `*_kink_free`: samples within 2e-5 of a kink are removed from the batch (tests/helpers.py::drop_fragile_points,
    < 1 % of the samples) -- the kernel must match the float32 oracle to the north-star tolerances (1e-5 / 1e-4);
  * the unfiltered batch: 1e-4 against the float32 oracle, or -- when the flips push the float32-vs-float32
    comparison beyond that -- adjudication by the SAME oracle in float64: the kernel must be no further from the
    float64 result than 1.5x the reference's own float32 path is.
Measured errors go to gpurun_out/r02_parity_sizes.json (copied to profiles/)."""
import json
import os

RESULTS = {}


def _record(name, **kv):
    RESULTS[name] = kv
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/r02_parity_sizes.json", "w") as fh:
            json.dump(RESULTS, fh, indent=1)
    except OSError:
        pass

import pytest
import torch

from helpers import drop_fragile_points, make_pair, oracle_like, oracle_mapping_chunked, rel_err
from miso_b200 import _lib, synth
from miso_b200.loss import MisoLossMapping

pytestmark = pytest.mark.gpu

TOL_TERM = 1e-5
TOL_GRAD = 1e-4


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _run_fused(net, mi, gt, poses, loss_type, w_eik, w_fs, trunc, eik_trunc):
    R, t = poses
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    for f in net.level_tensors():
        f.grad = None
    L = MisoLossMapping(loss_type=loss_type, weight_sdf=1.0, weight_eik=w_eik, weight_fs=w_fs, trunc_dist=trunc,
                        grad_method="autograd", eik_trunc_dist=eik_trunc)
    ld = L.compute(net, _cuda(mi), _cuda(gt))
    sum(v.mean() for v in ld.values()).backward()
    torch.cuda.synchronize()
    return {k: float(v) for k, v in ld.items()}


def _check(net, o2, got, want, name, rerun64=None, n_levels=2):
    """`rerun64(o64)` re-evaluates the oracle on the float64 copy of the model (adjudication, see module doc)."""
    assert set(got) == set(want)
    rec = {"terms_rel_err": {}, "grad_rel_err_vs_fp32_oracle": []}
    for k in want:
        rec["terms_rel_err"][k] = abs(got[k] - want[k]) / max(abs(want[k]), 1e-12)
    errs = [rel_err(net.features[l].feature.grad, o2.features[l].grad) for l in range(n_levels)]
    rec["grad_rel_err_vs_fp32_oracle"] = errs
    if max(errs) >= TOL_GRAD and rerun64 is not None:
        o64 = oracle_like(o2)
        rerun64(o64)
        ours = [rel_err(net.features[l].feature.grad, o64.features[l].grad) for l in range(n_levels)]
        ref32 = [rel_err(o2.features[l].grad, o64.features[l].grad) for l in range(n_levels)]
        rec["grad_rel_err_vs_fp64_oracle"] = ours
        rec["fp32_oracle_rel_err_vs_fp64_oracle"] = ref32
        _record(name, **rec)
        for l in range(n_levels):
            assert ours[l] < max(TOL_GRAD, 1.5 * ref32[l]), ("grid grad level vs fp64", l, ours[l], ref32[l])
    else:
        _record(name, **rec)
        for l in range(n_levels):
            assert errs[l] < TOL_GRAD, ("grid grad level", l, errs[l])
    for k in want:
        tol = TOL_GRAD if k == "eik" else TOL_TERM   # the eikonal term is a function of first-order gradients
        assert rec["terms_rel_err"][k] <= tol, (k, got[k], want[k])


@pytest.mark.parametrize("log2n", [18, 20])
def test_mapping_step_scannet_grid_full_batches(log2n):
    """BASELINE configs[0] (2^18) and configs[1] (2^20 points / iteration) on the ScanNet-submap grid."""
    bound = synth.SCANNET_SUBMAP_BOUND
    N = 1 << log2n
    net, _, o2 = make_pair(bound=bound, base_cell=0.5, scale=5, std=1e-2, seed=3, num_poses=49)
    mi, gt, poses = synth.rgbd_batch(N, num_kf=49, bound=bound, seed=55)
    assert _lib.load().miso_get_tuning(b"tc2_groups") == 4 and N // 128 > 3 * 148 * 4   # several tiles per group
    got = _run_fused(net, mi, gt, poses, "L1", 0.5, 0.1, 0.15, None)
    want = oracle_mapping_chunked(o2, mi, gt, poses, "L1", 1.0, 0.5, 0.1, 0.15, None)
    _check(net, o2, got, want, f"scannet_2p{log2n}",
           lambda o64: oracle_mapping_chunked(o64, mi, gt, poses, "L1", 1.0, 0.5, 0.1, 0.15, None))


@pytest.mark.parametrize("log2n", [18, 20])
def test_mapping_step_scannet_grid_kink_free(log2n):
    """Same sizes on the kink-free subset of a larger pool: strict north-star tolerances."""
    bound = synth.SCANNET_SUBMAP_BOUND
    N = 1 << log2n
    net, _, o2 = make_pair(bound=bound, base_cell=0.5, scale=5, std=1e-2, seed=3, num_poses=49)
    mi, gt, poses = synth.rgbd_batch(N + N // 16, num_kf=49, bound=bound, seed=56)
    mi, gt, dropped = drop_fragile_points(o2, mi, gt, poses, N, 0.15)
    got = _run_fused(net, mi, gt, poses, "L1", 0.5, 0.1, 0.15, None)
    want = oracle_mapping_chunked(o2, mi, gt, poses, "L1", 1.0, 0.5, 0.1, 0.15, None)
    _check(net, o2, got, want, f"scannet_2p{log2n}_kink_free")
    RESULTS[f"scannet_2p{log2n}_kink_free"]["samples_dropped_from_pool"] = dropped


def test_mapping_step_trainable_decoder_scannet_2p20_kink_free():
    """decoder.fix: False at BASELINE configs[1] size: every CTA of miso_mapping_step_wgrad walks ~55 tiles, the
    per-CTA rows are summed by the finalize kernel.  Decoder-parameter gradients vs the oracle's autograd."""
    bound = synth.SCANNET_SUBMAP_BOUND
    N = 1 << 20
    net, _, o2 = make_pair(bound=bound, base_cell=0.5, scale=5, std=1e-2, seed=3, num_poses=49, fix=False)
    mi, gt, poses = synth.rgbd_batch(N + N // 16, num_kf=49, bound=bound, seed=57)
    mi, gt, dropped = drop_fragile_points(o2, mi, gt, poses, N, 0.15)
    for p in net.decoder.parameters():
        p.grad = None
    got = _run_fused(net, mi, gt, poses, "L1", 0.5, 0.1, 0.15, None)
    want = oracle_mapping_chunked(o2, mi, gt, poses, "L1", 1.0, 0.5, 0.1, 0.15, None)
    _check(net, o2, got, want, "scannet_2p20_trainable_decoder_kink_free")
    errs = {}
    for name, a, b in zip(["W1", "b1", "W2", "b2", "W3", "b3"], net.decoder.parameters(), o2.decoder.parameters()):
        errs[name] = rel_err(a.grad, b.grad)
    RESULTS["scannet_2p20_trainable_decoder_kink_free"]["decoder_grad_rel_err_vs_fp32_oracle"] = errs
    RESULTS["scannet_2p20_trainable_decoder_kink_free"]["samples_dropped_from_pool"] = dropped
    _record("scannet_2p20_trainable_decoder_kink_free", **RESULTS["scannet_2p20_trainable_decoder_kink_free"])
    for name, e in errs.items():
        assert e < TOL_GRAD, (name, e)


def test_mapping_step_ncd_quad_grid_2p22_lidar_kink_free():
    bound = synth.NCD_QUAD_BOUND
    N = 1 << 22
    net, _, o2 = make_pair(bound=bound, base_cell=1.0, scale=5, std=1e-2, seed=5, num_poses=8)
    mi, gt, poses = synth.lidar_batch(N + N // 16, num_kf=8, seed=4)
    mi, gt, dropped = drop_fragile_points(o2, mi, gt, poses, N, 0.5)
    got = _run_fused(net, mi, gt, poses, "L2", 0.5, 0.5, 0.5, None)
    want = oracle_mapping_chunked(o2, mi, gt, poses, "L2", 1.0, 0.5, 0.5, 0.5, None, chunk=1 << 19)
    _check(net, o2, got, want, "ncd_quad_2p22_kink_free")
    RESULTS["ncd_quad_2p22_kink_free"]["samples_dropped_from_pool"] = dropped


def test_mapping_step_scannet_grid_eik_filter_and_ragged_tail():
    """2^18 + 77 points (last tile ragged, last group idle) with the |gt| < eik_trunc filter and L2."""
    bound = synth.SCANNET_SUBMAP_BOUND
    N = (1 << 18) + 77
    net, _, o2 = make_pair(bound=bound, base_cell=0.5, scale=5, std=1e-2, seed=4, num_poses=49)
    mi, gt, poses = synth.rgbd_batch(N, num_kf=49, bound=bound, seed=7)
    got = _run_fused(net, mi, gt, poses, "L2", 0.5, 0.5, 0.15, 0.1)
    want = oracle_mapping_chunked(o2, mi, gt, poses, "L2", 1.0, 0.5, 0.5, 0.15, 0.1)
    _check(net, o2, got, want, "scannet_2p18_ragged_eikfilter",
           lambda o64: oracle_mapping_chunked(o64, mi, gt, poses, "L2", 1.0, 0.5, 0.5, 0.15, 0.1))


def test_mapping_step_ncd_quad_grid_2p22_lidar():
    """BASELINE configs[3]: 2^22 LiDAR points on the NCD quad grid (fine level 324 MB, not L2 resident)."""
    bound = synth.NCD_QUAD_BOUND
    N = 1 << 22
    net, _, o2 = make_pair(bound=bound, base_cell=1.0, scale=5, std=1e-2, seed=5, num_poses=8)
    assert tuple(net.features[1].feature.shape) == (1, 4, 100, 450, 450)
    mi, gt, poses = synth.lidar_batch(N, num_kf=8, seed=3)
    got = _run_fused(net, mi, gt, poses, "L2", 0.5, 0.5, 0.5, None)
    want = oracle_mapping_chunked(o2, mi, gt, poses, "L2", 1.0, 0.5, 0.5, 0.5, None, chunk=1 << 19)
    _check(net, o2, got, want, "ncd_quad_2p22",
           lambda o64: oracle_mapping_chunked(o64, mi, gt, poses, "L2", 1.0, 0.5, 0.5, 0.5, None, chunk=1 << 19))


VARIANTS = [dict(tc2_groups=3), dict(pair=0), dict(tc2_groups=0), dict(mlp_tc=0), dict(force_int64=1),
            dict(tc2_groups=3, pair=0)]


@pytest.mark.parametrize("variant", VARIANTS, ids=lambda v: ",".join(f"{k}={x}" for k, x in v.items()))
def test_mapping_step_kernel_variants_multi_tile(variant):
    """MISO_TC2_GROUPS=3 / MISO_PAIR=0 / MISO_TC=1 / MISO_MLP=simt / forced 64-bit offsets at 2^18 points."""
    bound = synth.SCANNET_SUBMAP_BOUND
    N = 1 << 18
    net, _, o2 = make_pair(bound=bound, base_cell=0.5, scale=5, std=1e-2, seed=3, num_poses=49)
    mi, gt, poses = synth.rgbd_batch(N + N // 16, num_kf=49, bound=bound, seed=55)
    mi, gt, _ = drop_fragile_points(o2, mi, gt, poses, N, 0.15)
    with _lib.tuning(**variant):
        got = _run_fused(net, mi, gt, poses, "L1", 0.5, 0.1, 0.15, None)
    want = oracle_mapping_chunked(o2, mi, gt, poses, "L1", 1.0, 0.5, 0.1, 0.15, None)
    _check(net, o2, got, want, "variant_" + ",".join(f"{k}={x}" for k, x in variant.items()),
           lambda o64: oracle_mapping_chunked(o64, mi, gt, poses, "L1", 1.0, 0.5, 0.1, 0.15, None))


def test_mapping_step_grid_beyond_int32_offsets():
    """A fine level of 2^29 voxels x 4 channels = 2^31 elements (8.6 GB): `fits_int32` is false, element offsets
    of the touched corners exceed 2^31.  The oracle sees the z-crop [Z-40, Z) of both levels that the points
    (z in the top 30 coarse cells... of the bound) touch; outside that crop the product's gradient must be exactly zero."""
    from miso_b200.models import GridNet
    X, Y, Z = 512, 1024, 1024                      # fine level; coarse level = /4
    bound = [[0.0, 512.0], [0.0, 1024.0], [0.0, 1024.0]]
    cfg = synth.model_cfg(bound, base_cell_size=4.0, per_level_scale=4, num_poses=1)
    net = GridNet(cfg, device="cuda")
    assert tuple(net.features[1].feature.shape) == (1, 4, Z, Y, X)
    assert net.features[1].feature.numel() >= 2 ** 31
    zc_f, zc_c = 64, 16                            # crop depth in fine / coarse voxels: z in [960, 1024)
    g = torch.Generator().manual_seed(0)
    crops = [torch.randn(1, 4, zc_c, Y // 4, X // 4, generator=g) * 1e-2, torch.randn(1, 4, zc_f, Y, X, generator=g) * 1e-2]
    with torch.no_grad():
        net.features[0].feature[:, :, Z // 4 - zc_c:].copy_(crops[0].cuda())
        net.features[1].feature[:, :, Z - zc_f:].copy_(crops[1].cuda())
    sd = synth.decoder_weights(8, seed=0)
    net.decoder.load_state_dict(sd)
    from oracle import oracle as O
    dec = O.make_decoder(8)
    dec.load_state_dict({k.replace("network.", ""): v for k, v in sd.items()})
    crop_bound = [[0.0, 512.0], [0.0, 1024.0], [960.0, 1024.0]]
    o2 = O.OracleGridNet(crop_bound, crops, dec, second_order=True)
    # points: x,y anywhere (some outside), z in [970, 1030] so every touched corner lies in the crop or above the grid
    N = 50000
    x = torch.rand(N, 3, generator=g) * torch.tensor([540.0, 1060.0, 60.0]) + torch.tensor([-14.0, -18.0, 970.0])
    sdf = torch.randn(N, 1, generator=g) * 0.2
    mi = {"coords_frame": x[None], "sample_frame_ids": torch.zeros(1, N, 1, dtype=torch.long),
          "weights": torch.rand(1, N, 1, generator=g) + 0.5}
    gt = {"sdf": sdf[None], "sdf_valid": (sdf.abs() < 0.15)[None], "sdf_signs": torch.sign(sdf)[None] * (sdf.abs() >= 0.15)[None]}
    poses = (torch.eye(3)[None], torch.zeros(1, 3, 1))
    got = _run_fused(net, mi, gt, poses, "L1", 0.5, 0.1, 0.15, None)
    want = oracle_mapping_chunked(o2, mi, gt, poses, "L1", 1.0, 0.5, 0.1, 0.15, None)
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= TOL_TERM * max(abs(want[k]), 1e-12), (k, got[k], want[k])
    gf = net.features[1].feature.grad
    gc = net.features[0].feature.grad
    assert rel_err(gf[:, :, Z - zc_f:], o2.features[1].grad) < TOL_GRAD
    assert rel_err(gc[:, :, Z // 4 - zc_c:], o2.features[0].grad) < TOL_GRAD
    assert torch.count_nonzero(gf[:, :, :Z - zc_f]) == 0 and torch.count_nonzero(gc[:, :, :Z // 4 - zc_c]) == 0


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4]: interpolation fwd / bwd / double-bwd at sweep sizes; configs[2]: full-size alignment pair
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log2n,channels", [(20, 4), (22, 4), (20, 16)])
def test_interpolation_plugin_at_sweep_sizes_vs_aten(log2n, channels):
    """grid_sample_3d forward + backward (grid and coordinate gradients) on 2^20 / 2^22 points of a ScanNet-fine-level
    sized grid against ATen's own CUDA kernels on the same device (the reference's first-order path,
    grid_modules.py:89-94); the scatter is compared at 1e-4 (both sides accumulate with float32 atomics)."""
    import torch.nn.functional as F
    from miso_b200 import cuda_gridsample as cu
    N = 1 << log2n
    g = torch.Generator(device="cuda").manual_seed(log2n + channels)
    feat = (torch.randn(1, channels, 200, 100, 200, device="cuda", generator=g) * 1e-2)
    ours_in = feat.contiguous(memory_format=torch.channels_last_3d).clone().requires_grad_(True)
    aten_in = feat.clone().requires_grad_(True)
    xn = (torch.rand(1, N, 1, 1, 3, device="cuda", generator=g) * 2.1 - 1.05)
    go = torch.randn(1, channels, N, 1, 1, device="cuda", generator=g)
    res = []
    for fn, inp in ((cu.grid_sample_3d, ours_in), (F.grid_sample, aten_in)):
        x = xn.clone().requires_grad_(True)
        out = fn(inp, x, padding_mode="zeros", align_corners=False)
        out.backward(go)
        res.append((out.detach(), inp.grad, x.grad))
    assert rel_err(res[0][0], res[1][0]) < 1e-5
    assert rel_err(res[0][1], res[1][1]) < TOL_GRAD
    assert rel_err(res[0][2], res[1][2]) < TOL_GRAD
    _record(f"interp_2p{log2n}_C{channels}", fwd=rel_err(res[0][0], res[1][0]), grad_input=rel_err(res[0][1], res[1][1]),
            grad_grid=rel_err(res[0][2], res[1][2]))


def test_double_backward_at_sweep_size_vs_reference_extension():
    """The eikonal pattern (gg_grid given) on 2^20 points: all three double-backward outputs against the reference's own
    `grid_sampler_3d_grad2_kernel` (oracle/_ref, built unmodified for sm_100a)."""
    from oracle import ref_gpu
    if not ref_gpu.available():
        pytest.skip("oracle/_ref/gridsample_grad2.so not built")
    from miso_b200 import cuda_gridsample as cu
    N = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(5)
    feat = torch.randn(1, 4, 100, 50, 100, device="cuda", generator=g) * 1e-1
    xn = torch.rand(1, N, 1, 1, 3, device="cuda", generator=g) * 2.1 - 1.05
    go = torch.randn(1, 4, N, 1, 1, device="cuda", generator=g)
    g2 = torch.randn(1, N, 1, 1, 3, device="cuda", generator=g)
    res = []
    for fn, inp0 in ((cu.grid_sample_3d, feat.contiguous(memory_format=torch.channels_last_3d)), (ref_gpu.grid_sample_3d, feat)):
        inp = inp0.clone().requires_grad_(True)
        x = xn.clone().requires_grad_(True)
        gor = go.clone().requires_grad_(True)
        out = fn(inp, x, padding_mode="zeros", align_corners=False)
        (gx,) = torch.autograd.grad(out, x, gor, create_graph=True)
        (gx * g2).sum().backward()
        res.append((gx.detach(), inp.grad, x.grad, gor.grad))
    for name, a, b in zip(("grad_grid", "dbl.g_input", "dbl.g_grid", "dbl.gg_output"), res[0], res[1]):
        assert rel_err(a, b) < TOL_GRAD, (name, rel_err(a, b))


def test_alignment_pair_at_full_scannet_size():
    """BASELINE configs[2] sizes for ONE overlapping pair of ScanNet-shaped submaps: level 0 (32 k samples) and level 1
    (4 M samples, 8 channels) of pairwise_loss_latent against the oracle: loss 1e-5, pose gradients 1e-4, in-bound
    count and intersection verdict exact."""
    from miso_b200.align import AlignBatch, pairwise_loss_latent
    from miso_b200.models import GridAtlas
    from oracle import oracle as O
    bound = synth.SCANNET_SUBMAP_BOUND
    Rt, tt = synth.submap_layout(2, spacing=(12.0, 12.0))
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=10.0, trans_m=0.5)
    atlas = GridAtlas(synth.model_cfg(bound, num_poses=1), device="cuda")
    subs = []
    shapes = O.level_shapes(bound, 0.5, 5, 2, 4)
    for i in range(2):
        atlas.add_submap(torch.tensor(bound), Rp[i], tp[i])
        feats = synth.fill_submap_from_field(shapes, bound, Rt[i], tt[i])
        with torch.no_grad():
            for l in range(2):
                atlas.get_submap(i).features[l].feature.copy_(feats[l].cuda())
        atlas.get_submap(i).lock_feature()
        subs.append(O.OracleGridNet(bound, feats, None))
    oat = O.OracleAtlas(subs, Rp, tp)
    atlas.precompute_coordinates_for_alignment()
    oat.precompute([0, 1])
    for level in (0, 1):
        assert torch.equal(atlas.coordinates_for_alignment(0, level).cpu(), oat.coords[(0, level)])
        for p in list(atlas.rotation_corrections) + list(atlas.translation_corrections):
            p.grad = None
        for q in oat.rot + oat.tra:
            q.grad = None
        (_, val), = pairwise_loss_latent(atlas, None, 0, 1, level=level, device="cuda").items()
        val.backward()
        Rs, ts = oat.updated_submap_pose(0)
        Rd, td = oat.updated_submap_pose(1)
        lo, mask, _ = O.pairwise_loss_latent(subs[0], subs[1], oat.coords[(0, level)], Rs, ts, Rd, td, level, return_aux=True)
        lo.backward()
        batch = AlignBatch(atlas, [(0, 1)], level, check_intersection=True, want_masks=True, cache_src_features=False)
        out = batch.launch(batch.pair_poses())
        assert int(out[0, 1].item()) == int(mask.sum())                         # in-bound count, exact
        assert torch.equal(batch.masks[0].bool().cpu(), mask[:, 0])
        assert rel_err(val, lo) < 1e-5, (level, rel_err(val, lo))
        for i in range(2):
            assert rel_err(atlas.rotation_corrections[i].grad, oat.rot[i].grad) < TOL_GRAD, (level, "rot", i)
            assert rel_err(atlas.translation_corrections[i].grad, oat.tra[i].grad) < TOL_GRAD, (level, "tra", i)
        _record(f"align_pair_level{level}", samples=int(oat.coords[(0, level)].shape[0]), valid=int(mask.sum()),
                loss_rel_err=rel_err(val, lo))
    batch = AlignBatch(atlas, [(0, 1)], 1, check_intersection=True)
    batch.update_intersections(batch.pair_poses())
    Rs, ts = oat.updated_submap_pose(0)
    Rd, td = oat.updated_submap_pose(1)
    assert bool(batch.enabled[0].item()) == bool(O.check_submap_intersection(subs[0], subs[1], Rs, ts, Rd, td))
