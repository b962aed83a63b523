"""Host logic of the slab-sharded fit (miso_b200/sharded_fit.py) on CPU tensors: slab cuts from a histogram, the
two-phase cost model of `calibrate` (max samples + c * max parameters over the ranks), the ownership rule."""
import torch

from miso_b200 import sharded_fit as sf
from miso_b200 import synth
from miso_b200.loss import MisoLossMapping
from miso_b200.models import GridNet

BOUND = [[-4.0, 4.0], [-2.0, 2.0], [-4.0, 4.0]]


def test_slab_bounds_from_histogram_balances_and_keeps_a_plane_per_rank():
    hist = torch.tensor([0, 0, 100, 0, 0, 0, 100, 0, 0, 0], dtype=torch.float64)
    b = sf.slab_bounds_from_histogram(hist, 2)
    assert b[0] == 0 and b[-1] == 10 and 3 <= b[1] <= 6            # the cut separates the two heavy planes
    b = sf.slab_bounds_from_histogram(torch.tensor([1000.0, 0, 0, 0]), 4)
    assert b == [0, 1, 2, 3, 4]                                      # every rank owns at least one plane
    try:
        sf.slab_bounds_from_histogram(torch.ones(3), 4)
    except ValueError:
        pass
    else:
        raise AssertionError("more ranks than planes must raise")


def test_plane_of_points_matches_the_kernels_index_arithmetic():
    Z, lo, hi = 40, -4.0, 4.0
    z = torch.tensor([-4.0, -3.9, 0.0, 3.99, 4.0, 7.0, float("nan"), -9.0])
    got = sf.plane_of_points(z, lo, hi, Z)
    iz = ((2 * (z - lo) / (hi - lo) - 1 + 1) * Z - 1) / 2
    want = torch.nan_to_num(torch.floor(iz), nan=0.0).clamp(0, Z - 1).long()
    assert torch.equal(got, want)
    assert got[0] == 0 and got[4] == Z - 1 and got[5] == Z - 1 and got[6] == 0 and got[7] == 0


def _model_and_batch(n=20000, skew=False):
    cfg = synth.model_cfg(BOUND, n_levels=2, feature_dim=4, base_cell_size=1.0, per_level_scale=4, num_poses=1)
    net = GridNet(cfg, device="cpu")
    net.set_initial_kf_pose(0, torch.eye(3), torch.zeros(3, 1), kf_key="KF0")
    net.lock_pose()
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(n, 3, generator=g) * 2 - 1) * torch.tensor([3.9, 1.9, 3.9])
    if skew:     # 60 % of the samples in a thin z-layer (an outdoor LiDAR batch: most returns near the ground)
        k = int(0.6 * n)
        x[:k, 2] = -3.5 + 0.2 * torch.rand(k, generator=g)
    mi = {"coords_frame": x[None], "sample_frame_ids": torch.zeros(1, n, 1, dtype=torch.long), "weights": torch.ones(1, n, 1)}
    return net, mi


def _modelled_step(fit, mi, bounds, axis, voxel_cost):
    f = fit.model.level_tensors()[fit.slab_level]
    n_planes = f.shape[2] if axis == 2 else f.shape[3]
    lo, hi = fit.model._bound_host[2 * axis], fit.model._bound_host[2 * axis + 1]
    plane = sf.plane_of_points(mi["coords_frame"][0][:, axis], lo, hi, n_planes)
    samples = torch.bincount(plane, minlength=n_planes).double()
    per_plane = f.numel() // n_planes
    return max(float(samples[a:b].sum()) for a, b in zip(bounds[:-1], bounds[1:])) \
        + voxel_cost * per_plane * max(b - a for a, b in zip(bounds[:-1], bounds[1:]))


def test_calibrate_single_rank_keeps_the_native_layout():
    net, mi = _model_and_batch()
    fit = sf.SlabShardedFit(net, MisoLossMapping(loss_type="L2", weight_eik=0.0), rank=0, world=1)
    assert fit.calibrate(mi) == [0, fit.Z] and fit.axis == 2 and not fit.p2p


def test_calibrate_picks_the_cheaper_axis_for_a_skewed_batch():
    net, mi = _model_and_batch(skew=True)
    loss = MisoLossMapping(loss_type="L2", weight_eik=0.0)
    vc = 1.0 / 56.0
    fits = [sf.SlabShardedFit(net, loss, rank=r, world=4) for r in range(4)]
    bounds = fits[0].calibrate(mi, voxel_cost=vc)
    axis = fits[0].axis
    assert axis == 1, "60 % of the samples sit in one thin z-layer: y-slabs balance, z-slabs cannot"
    # the level is stored with the slab axis slowest, values and logical shape unchanged
    p = net.level_tensors()[fits[0].slab_level]
    assert p.shape[3] == fits[0].Z and p.stride(3) > p.stride(2) > p.stride(4) > p.stride(1) == 1
    # no other candidate of the search (both axes x the four weightings) has a smaller modelled step
    f = net.level_tensors()[fits[0].slab_level]
    best = _modelled_step(fits[0], mi, bounds, axis, vc)
    for ax in (2, 1):
        n_planes = f.shape[2] if ax == 2 else f.shape[3]
        lo, hi = net._bound_host[2 * ax], net._bound_host[2 * ax + 1]
        samples = torch.bincount(sf.plane_of_points(mi["coords_frame"][0][:, ax], lo, hi, n_planes), minlength=n_planes).double()
        for w in (1.0, 0.5, 0.25, 0.0):
            cand = sf.slab_bounds_from_histogram(samples + w * vc * (f.numel() // n_planes), 4)
            assert best <= _modelled_step(fits[0], mi, cand, ax, vc) / 0.98 + 1e-9
    # every rank of the same batch derives the same cut; slabs tile the level
    assert bounds[0] == 0 and bounds[-1] == fits[0].Z and all(b > a for a, b in zip(bounds[:-1], bounds[1:]))
    fits[0].restore_layout()
    for r in range(1, 4):
        assert fits[r].calibrate(mi, voxel_cost=vc) == bounds
        fits[r].restore_layout()
    assert fits[0].axis == 2 and net.level_tensors()[fits[0].slab_level].stride(2) > net.level_tensors()[fits[0].slab_level].stride(3)


def test_replicated_gradients_and_loss_terms_share_one_buffer():
    """p2p mode sums the replicated levels' gradients and the 4 loss terms with ONE all_reduce: their storage is one
    flat buffer, each .grad a strided view of it with the parameter's own layout, existing gradient values kept."""
    net, _ = _model_and_batch(n=16)
    fit = sf.SlabShardedFit(net, MisoLossMapping(loss_type="L2", weight_eik=0.0), rank=0, world=1)
    feats = net.level_tensors()
    coarse = [f for l, f in enumerate(feats) if l != fit.slab_level][0]
    coarse.grad = torch.arange(coarse.numel(), dtype=torch.float32).reshape(coarse.shape).contiguous(
        memory_format=torch.channels_last_3d) if coarse.stride(1) == 1 else torch.ones_like(coarse)
    before = coarse.grad.clone()
    terms = fit._pack_replicated_grads(feats)
    key, flat, total = fit._flat_buf
    assert terms.numel() == 4 and total == coarse.numel() and flat.numel() == total + 4
    assert terms.data_ptr() == flat.data_ptr() + 4 * total
    assert coarse.grad.data_ptr() == flat.data_ptr() and coarse.grad.stride() == coarse.stride()
    assert torch.equal(coarse.grad, before)
    coarse.grad.add_(1.0)
    terms.fill_(7.0)
    assert float(flat[:total].sum()) == float((before + 1.0).sum()) and flat[total:].tolist() == [7.0] * 4
    assert fit._pack_replicated_grads(feats).data_ptr() == terms.data_ptr()      # cached: same buffer on the next step
    assert feats[fit.slab_level].grad is None                                     # the slab level is not replicated
