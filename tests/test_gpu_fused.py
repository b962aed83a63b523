"""GPU parity of the fused field kernels (C-ABI section 2) against the CPU oracle: features, SDF,
analytic gradient, first-order backward, eikonal double-backward, the one-kernel mapping step, Adam.
Tolerances (BASELINE.json north_star): forward 1e-5 relative; first/second-order gradients 1e-4."""
import numpy as np
import pytest
import torch

from helpers import SMALL_BOUND, make_pair, points_in, rel_err
from miso_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL_F = 1e-5
TOL_G = 1e-4


@pytest.mark.parametrize("n_levels,fdim", [(2, 4), (1, 4), (3, 4), (2, 8), (1, 16)])
def test_query_feature_and_forward(n_levels, fdim):
    net, o1, _ = make_pair(n_levels=n_levels, fdim=fdim, scale=3)
    x = points_in(SMALL_BOUND, 5000, seed=1)
    with torch.no_grad():
        f = net.query_feature(x.cuda())
        y = net(x.cuda())
    assert f.shape == (5000, n_levels * fdim)
    assert rel_err(f, o1.query_feature(x)) < TOL_F
    assert rel_err(y, o1(x)) < TOL_F
    assert y.shape == (5000, 1)


def test_forward_ignore_level_and_empty():
    net, o1, _ = make_pair()
    net.ignore_level(1)
    o1.ignore_level_[1] = True
    x = points_in(SMALL_BOUND, 1000, seed=2)
    with torch.no_grad():
        assert rel_err(net(x.cuda()), o1(x)) < TOL_F
        assert rel_err(net.query_feature(x.cuda()), o1.query_feature(x)) < TOL_F
        assert net(torch.zeros(0, 3, device="cuda")).shape == (0, 1)


def test_first_order_backward_and_gradx():
    net, o1, o2 = make_pair()
    net.unlock_feature()
    x = points_in(SMALL_BOUND, 4000, seed=3)
    w = torch.randn(4000, 1, generator=torch.Generator().manual_seed(4))
    xg = x.cuda().requires_grad_(True)
    y = net(xg)
    (y * w.cuda()).sum().backward()
    xo = x.clone().requires_grad_(True)
    yo = o1(xo)
    (yo * w).sum().backward()
    assert rel_err(y, yo) < TOL_F
    assert rel_err(xg.grad, xo.grad) < TOL_G
    for l in range(2):
        assert rel_err(net.features[l].feature.grad, o1.features[l].grad) < TOL_G
    # analytic spatial gradient output == gradient3d(..., 'autograd')
    from miso_b200.diff import gradient3d
    g = gradient3d(x.cuda().requires_grad_(True), net, method="autograd", create_graph=False)
    go = O.gradient3d(x.clone().requires_grad_(True), o2, method="autograd", create_graph=False)
    assert rel_err(g, go) < TOL_G


@pytest.mark.parametrize("via", ["forward_with_gradient", "autograd.grad"])
def test_eikonal_double_backward(via):
    """mean((|grad_x f| - 1)^2) -> d/d grid : the path that needs grid_sampler_3d_grad2_kernel in the reference."""
    net, _, o2 = make_pair()
    net.unlock_feature()
    x = points_in(SMALL_BOUND, 3000, seed=5, scale=0.95)
    xg = x.cuda().requires_grad_(True)
    if via == "forward_with_gradient":
        _, g = net.forward_with_gradient(xg)
    else:
        y = net(xg)
        g = torch.autograd.grad(y, xg, torch.ones_like(y), create_graph=True)[0]
    loss = torch.mean((g.norm(dim=-1) - 1) ** 2)
    loss.backward()
    xo = x.clone().requires_grad_(True)
    go = O.gradient3d(xo, o2, "autograd", create_graph=True)
    lo = torch.mean((go.norm(dim=-1) - 1) ** 2)
    lo.backward()
    assert rel_err(loss, lo) < TOL_G
    for l in range(2):
        assert rel_err(net.features[l].feature.grad, o2.features[l].grad) < TOL_G
    # coordinate cotangent of the double backward (mixed second derivatives)
    assert rel_err(xg.grad, xo.grad) < TOL_G


def _batch(n=6000, num_kf=4, seed=1):
    return synth.rgbd_batch(n, num_kf=num_kf, bound=SMALL_BOUND, seed=seed, wall_margin=0.3)


def _to_cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("loss_type,w_eik,eik_trunc,w_fs", [("L1", 0.0, None, 0.1), ("L1", 0.5, None, 0.1),
                                                             ("L2", 0.5, 0.1, 0.5), ("L2", 0.0, None, 0.0)])
def test_mapping_step_fused_vs_oracle(loss_type, w_eik, eik_trunc, w_fs):
    """MisoLossMapping.compute -> backward: every term and d(total)/d(grid) against the oracle's
    restatement of loss.py:754-813 (autograd second-order eikonal through the gather oracle)."""
    from miso_b200.loss import MisoLossMapping
    net, _, o2 = make_pair()
    mi, gt, (R, t) = _batch()
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    L = MisoLossMapping(loss_type=loss_type, weight_sdf=1.0, weight_eik=w_eik, weight_fs=w_fs, trunc_dist=0.15,
                        grad_method="autograd", eik_trunc_dist=eik_trunc)
    ld = L.compute(net, _to_cuda(mi), _to_cuda(gt))
    total = sum(v.mean() for v in ld.values())
    total.backward()
    lo = O.mapping_loss(o2, mi, gt, {k: (R[k], t[k]) for k in range(R.shape[0])}, loss_type, 1.0, w_eik, w_fs, 0.15,
                        grad_method="autograd", eik_trunc_dist=eik_trunc)
    sum(lo.values()).backward()
    assert set(ld.keys()) == set(lo.keys())
    for k in lo:
        assert rel_err(ld[k], lo[k]) < TOL_G, k
    for l in range(2):
        assert rel_err(net.features[l].feature.grad, o2.features[l].grad) < TOL_G


def test_mapping_generic_path_matches_fused_and_oracle():
    """Trainable decoder -> generic path (twice-differentiable per-level op + torch MLP), incl. decoder grads
    and the finite-difference eikonal (configured default grad_method, scannet.yaml:49)."""
    from miso_b200.loss import MisoLossMapping
    net, o1, _ = make_pair(fix=False)
    assert net.fused_spec() is None
    mi, gt, (R, t) = _batch(3000)
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    L = MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                        grad_method="finitediff", finite_diff_eps=0.024, eik_trunc_dist=None)
    ld = L.compute(net, _to_cuda(mi), _to_cuda(gt))
    sum(v.mean() for v in ld.values()).backward()
    lo = O.mapping_loss(o1, mi, gt, {k: (R[k], t[k]) for k in range(R.shape[0])}, "L1", 1.0, 0.5, 0.1, 0.15,
                        finite_diff_eps=0.024, grad_method="finitediff", eik_trunc_dist=None)
    sum(lo.values()).backward()
    for k in lo:
        assert rel_err(ld[k], lo[k]) < TOL_G, k
    for l in range(2):
        assert rel_err(net.features[l].feature.grad, o1.features[l].grad) < TOL_G
    gp = [p.grad for p in net.decoder.parameters()]
    op = [p.grad for p in o1.decoder.parameters()]
    for a, b in zip(gp, op):
        assert rel_err(a, b) < TOL_G


@pytest.mark.parametrize("loss_type,w_eik,eik_trunc,w_fs,shape,n", [
    ("L1", 0.5, None, 0.1, (2, 4), 6000), ("L2", 0.5, 0.1, 0.5, (2, 4), 4133), ("L1", 0.0, None, 0.1, (2, 4), 3000),
    ("L1", 0.5, None, 0.1, (1, 16), 3000), ("L2", 0.3, None, 0.0, (3, 4), 2500), ("L1", 0.5, None, 0.1, (1, 4), 64)])
def test_mapping_trainable_decoder_fused(loss_type, w_eik, eik_trunc, w_fs, shape, n):
    """decoder.fix: False (grid_net.py:110,126,346-348) stays on the fused step: the grid gradients and loss terms come
    from miso_mapping_step, d total / d {W1,b1,W2,b2,W3,b3} -- including the eikonal term's second-order path through
    the weights -- from miso_mapping_step_wgrad; both against the oracle's autograd on a kink-free batch."""
    from helpers import drop_fragile_points
    from miso_b200.loss import MisoLossMapping
    net, _, o2 = make_pair(n_levels=shape[0], fdim=shape[1], fix=False, scale=3 if shape[0] > 2 else 5)
    mi, gt, (R, t) = _batch(n + 400, seed=7)
    mi, gt, _ = drop_fragile_points(o2, mi, gt, (R, t), n, 0.15)
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    L = MisoLossMapping(loss_type=loss_type, weight_sdf=1.0, weight_eik=w_eik, weight_fs=w_fs, trunc_dist=0.15,
                        grad_method="autograd", eik_trunc_dist=eik_trunc)
    assert any(p.requires_grad for p in net.decoder.parameters()) and net.fused_spec() is None
    assert L._fused_ok(net)
    ld = L.compute(net, _to_cuda(mi), _to_cuda(gt))
    sum(v.mean() for v in ld.values()).backward()
    lo = O.mapping_loss(o2, mi, gt, {k: (R[k], t[k]) for k in range(R.shape[0])}, loss_type, 1.0, w_eik, w_fs, 0.15,
                        grad_method="autograd", eik_trunc_dist=eik_trunc)
    sum(lo.values()).backward()
    for k in lo:
        assert rel_err(ld[k], lo[k]) < TOL_G, k
    for l in range(shape[0]):
        assert rel_err(net.features[l].feature.grad, o2.features[l].grad) < TOL_G
    names = ["W1", "b1", "W2", "b2", "W3", "b3"]
    got = [p.grad.clone() for p in net.decoder.parameters()]
    for name, a, b in zip(names, got, [p.grad for p in o2.decoder.parameters()]):
        assert a.shape == b.shape
        assert rel_err(a, b) < TOL_G, name
    # the trainer's entry point accumulates the same gradients into .grad (twice the value after a second pass)
    L.step_into_grads(net, _to_cuda(mi), _to_cuda(gt))
    for name, a, p in zip(names, got, net.decoder.parameters()):
        assert rel_err(p.grad, 2 * a) < 1e-6, name


def test_trainer_with_trainable_decoder_matches_oracle_adam():
    """Three fused train steps with decoder.fix: False vs oracle loss + torch.optim.Adam over grids AND decoder."""
    from helpers import drop_fragile_points
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import GridTrainer
    net, _, o2 = make_pair(fix=False)
    mi, gt, (R, t) = _batch(4400, seed=9)
    mi, gt, _ = drop_fragile_points(o2, mi, gt, (R, t), 4000, 0.15)
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    L = MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                        grad_method="autograd", eik_trunc_dist=None)
    tr = GridTrainer({"grid_training_mode": "joint", "learning_rate": 1e-3}, net, L, lambda e: (mi, gt))
    opt = torch.optim.Adam(list(o2.features.parameters()) + list(o2.decoder.parameters()), lr=1e-3)
    poses = {k: (R[k], t[k]) for k in range(R.shape[0])}
    for _ in range(3):
        terms = tr.train_step(_to_cuda(mi), _to_cuda(gt))
        opt.zero_grad()
        lo = O.mapping_loss(o2, mi, gt, poses, "L1", 1.0, 0.5, 0.1, 0.15, grad_method="autograd", eik_trunc_dist=None)
        tot = sum(lo.values())
        tot.backward()
        opt.step()
        assert rel_err(terms[3], tot) < TOL_G
    # Adam's first steps are sign-like (|update| ~ lr regardless of the gradient's size): compare the parameters
    for a, b, b0 in zip(net.decoder.parameters(), o2.decoder.parameters(), synth.decoder_weights(8, seed=0).values()):
        assert rel_err(a, b) < 2e-4
        if b0.numel() > 1:
            assert rel_err(b, b0) > 1e-3     # the decoder did move
    for l in range(2):
        assert rel_err(net.features[l].feature, o2.features[l]) < 1e-3


def test_trainer_steps_match_oracle_adam():
    """k fused train steps (count + mapping step + fused Adam/zero) vs oracle loss + torch.optim.Adam."""
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import GridTrainer
    net, _, o2 = make_pair()
    mi, gt, (R, t) = _batch(4000)
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    L = MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                        grad_method="autograd", eik_trunc_dist=None)
    tr = GridTrainer({"epochs": 5, "learning_rate": 1e-3, "grid_training_mode": "joint"}, net, L,
                     lambda e: (mi, gt), device="cuda")
    losses = [tr.train_epoch(e) for e in range(5)]
    opt = torch.optim.Adam(list(o2.features.parameters()), lr=1e-3)
    ol = []
    for e in range(5):
        opt.zero_grad()
        lo = O.mapping_loss(o2, mi, gt, {k: (R[k], t[k]) for k in range(R.shape[0])}, "L1", 1.0, 0.5, 0.1, 0.15,
                            grad_method="autograd", eik_trunc_dist=None)
        tot = sum(lo.values())
        tot.backward()
        opt.step()
        ol.append(float(tot))
    got = [float(l[3]) for l in losses]
    assert np.allclose(got, ol, rtol=1e-4), (got, ol)
    for l in range(2):
        assert rel_err(net.features[l].feature, o2.features[l]) < TOL_G
        # the fused Adam leaves a zeroed gradient buffer behind
        assert torch.count_nonzero(net.features[l].feature.grad) == 0


def test_adam_matches_torch():
    from miso_b200.optim import FusedAdam
    torch.manual_seed(0)
    p = torch.randn(1, 4, 6, 5, 7, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    p1 = torch.nn.Parameter(p.clone())
    p2 = torch.nn.Parameter(p.clone())
    o1 = FusedAdam([p1], lr=1e-3)
    o2 = torch.optim.Adam([p2], lr=1e-3)
    for i in range(4):
        g = torch.randn_like(p)
        p1.grad = g.clone()
        p2.grad = g.clone()
        o1.step()
        o2.step()
    assert rel_err(p1, p2) < 1e-6
    assert torch.count_nonzero(p1.grad) == 0
    # sparse gradients (most voxels never touched, some touched once and then only decaying): the bitmap-tracked
    # kernel, the plain kernel and torch.optim.Adam must agree, and the bitmap must mark exactly the touched voxels
    base = torch.randn(1, 4, 16, 12, 20, device="cuda").contiguous(memory_format=torch.channels_last_3d)
    ps = [torch.nn.Parameter(base.clone()) for _ in range(3)]
    opts = [FusedAdam([ps[0]], lr=1e-3, track_touched=True), FusedAdam([ps[1]], lr=1e-3, track_touched=False),
            torch.optim.Adam([ps[2]], lr=1e-3)]
    ever = torch.zeros(16, 12, 20, dtype=torch.bool, device="cuda")
    for i in range(5):
        mask = (torch.rand(16, 12, 20, device="cuda") < 0.1) if i < 3 else torch.zeros_like(ever)
        ever |= mask
        g = (torch.randn_like(base) * mask[None, None]).contiguous(memory_format=torch.channels_last_3d)
        for q, o in zip(ps, opts):
            q.grad = g.clone()
            o.step()
    assert rel_err(ps[0], ps[2]) < 1e-6 and rel_err(ps[1], ps[2]) < 1e-6
    assert torch.equal(ps[0].detach(), ps[1].detach())                      # bit-identical to the plain kernel
    words = opts[0].state[ps[0]]["touched"]
    bits = ((words.view(-1, 1) >> torch.arange(32, device="cuda")) & 1).bool().reshape(-1)[:ever.numel()]
    assert torch.equal(bits, ever.reshape(-1))                               # channels-last: voxel order = z,y,x


def test_full_size_properties():
    """Size-independent checks at BASELINE config-1 scale (2^18 points, ScanNet-submap grid):
    linearity of the interpolation in the grid values and zero output outside the bound."""
    from miso_b200 import field
    bound = synth.SCANNET_SUBMAP_BOUND
    net, _, _ = make_pair(bound=bound, base_cell=0.5, scale=5, seed=3)
    x = points_in(bound, 2 ** 18, seed=9).cuda()
    feats = net.level_tensors()
    b = net._bound_host
    with torch.no_grad():
        f1 = field.field_features_raw(feats, b, x)
        f2 = field.field_features_raw([2.0 * f for f in feats], b, x)
        assert rel_err(f2, 2.0 * f1) < 1e-6
        far = x + 100.0
        assert torch.count_nonzero(field.field_features_raw(feats, b, far)) == 0
        # interpolating a constant field returns the constant strictly inside the grid
        ones = [torch.ones_like(f) for f in feats]
        inside = points_in(bound, 2 ** 16, seed=10, scale=0.9).cuda()
        assert rel_err(field.field_features_raw(ones, b, inside), torch.ones(2 ** 16, 8)) < 1e-6


def test_tracker_lm_normal_equations_and_step():
    """Tracker.lm_step (tracker.py:148-212): fused sdf + analytic gradient -> J^T W J, J^T W r, 6x6 solve,
    against the oracle's restatement (autograd gradient)."""
    from miso_b200.tracker import Tracker
    net, _, o2 = make_pair()
    g = torch.Generator().manual_seed(0)
    xf = (torch.rand(4000, 3, generator=g) - 0.5) * torch.tensor([2.0, 1.0, 2.0])
    gt_sdf = torch.randn(4000, 1, generator=g) * 0.05
    w0 = torch.tensor([[0.1, -0.2, 0.05]])
    Rwf = O.so3_exp_map(w0)[0]
    twf = torch.tensor([[0.1], [0.05], [-0.1]])
    net.set_initial_kf_pose(0, Rwf, twf, kf_key="KF0")
    for loss_type in ("L2", "GM"):
        tr = Tracker(net, loss_type=loss_type, gm_scale_sdf=0.1, lm_lambda=1e-4)
        H, b, fov = tr.normal_equations(xf.cuda(), gt_sdf.cuda(), Rwf.cuda(), twf.cuda())
        Ho, bo, do = O.lm_normal_equations(o2, xf, gt_sdf, Rwf, twf, loss_type=loss_type, gm_scale=0.1, lm_lambda=1e-4)
        assert rel_err(H, Ho) < TOL_G and rel_err(b, bo) < TOL_G
        assert rel_err(torch.linalg.solve(H, -b), do) < TOL_G   # measured 3.6e-7
    mi = {"coords_frame": xf[None].cuda(), "sample_frame_ids": torch.zeros(1, 4000, 1, dtype=torch.long).cuda()}
    gt = {"sdf": gt_sdf[None].cuda(), "sdf_valid": torch.ones(1, 4000, 1, dtype=torch.bool).cuda()}
    info = tr.lm_step(0, mi, gt)
    assert rel_err(net.rotation_corrections[0], do[:3, 0]) < TOL_G
    assert rel_err(net.translation_corrections[0], do[3:]) < TOL_G
    assert 0.0 <= info["fov_overlap"] <= 1.0
    # one-launch normal equations vs the torch formulation on the same fused sdf+gradient, incl. the in-kernel
    # |gt| < trunc filter against an explicit nonzero()/gather selection (tracker.py:158-164), and fov_overlap
    for loss_type in ("L2", "GM"):
        tr = Tracker(net, loss_type=loss_type, gm_scale_sdf=0.1, lm_lambda=1e-4)
        sel = torch.nonzero(torch.abs(gt_sdf[:, 0]) < 0.04, as_tuple=False).squeeze(1)
        H1, b1, f1 = tr.normal_equations(xf.cuda(), gt_sdf.cuda(), Rwf.cuda(), twf.cuda(), trunc_dist=0.04)
        H2, b2, f2 = tr.normal_equations_torch(xf[sel].cuda(), gt_sdf[sel].cuda(), Rwf.cuda(), twf.cuda())
        assert rel_err(H1, H2) < TOL_G and rel_err(b1, b2) < TOL_G
        assert abs(float(f1) - float(f2)) < 1e-6


def test_inference_forward_and_atlas_query():
    """Forward-only path (no Jacobian pass) and GridAtlas.query_feature's masked mean (grid_atlas.py:374-391)."""
    from miso_b200.models import GridAtlas
    net, o1, _ = make_pair()
    x = points_in(SMALL_BOUND, 20000, seed=11)
    with torch.no_grad():
        y = net(x.cuda())
    assert rel_err(y, o1(x)) < TOL_F
    atlas = GridAtlas(synth.model_cfg(SMALL_BOUND, num_poses=1), device="cuda")
    ors = []
    for i in range(2):
        R = O.so3_exp_map(torch.tensor([[0.0, 0.1 * i, 0.0]]))[0]
        t = torch.tensor([[1.0 * i], [0.0], [0.5 * i]])
        atlas.add_submap(torch.tensor(SMALL_BOUND), R, t)
        gsm = torch.Generator().manual_seed(i)
        feats = [torch.randn(f.feature.shape, generator=gsm) * 0.1 for f in atlas.get_submap(i).features]
        with torch.no_grad():
            for l in range(2):
                atlas.get_submap(i).features[l].feature.copy_(feats[l].cuda())
        ors.append((O.OracleGridNet(SMALL_BOUND, feats, None), R, t))
    xw = points_in(SMALL_BOUND, 5000, seed=12, scale=1.3)
    with torch.no_grad():
        got = atlas.query_feature(xw.cuda())
    sf, sw = 0, 0
    for om, R, t in ors:
        xs = O.transfrom_points_from(xw, R, t)
        m = O.coords_in_bound(xs, om.bound)
        sf = sf + m * om.query_feature(xs)
        sw = sw + m
    sw = torch.where(sw == 0, torch.ones_like(sw), sw).float()
    assert rel_err(got, sf / sw) < TOL_F
    # the one-launch atlas query (miso_atlas_features, taken under no_grad) vs the per-submap torch path (taken
    # when autograd is needed), incl. a three-submap atlas with one inactive submap
    xg = xw.cuda().requires_grad_(True)
    per_submap = atlas.query_feature(xg)
    assert rel_err(got, per_submap) < TOL_F
    atlas.add_submap(torch.tensor(SMALL_BOUND), torch.eye(3), torch.tensor([[-1.0], [0.2], [0.0]]))
    atlas.get_submap(2).randn_features(0.1)
    atlas.active_submaps = [0, 2]
    with torch.no_grad():
        fused = atlas.query_feature(xw.cuda())
    assert rel_err(fused, atlas.query_feature(xg)) < TOL_F


def test_compact_host_batches_match_reference_format():
    """CompactBatch (int16 ids, masks rebuilt by miso_expand_batch) through train_host_batches gives the same losses
    and the same parameters as the reference-format batches -- bit-exact masks/ids, so only atomics order differs --
    and refuses batches whose masks are not the dataset's functions of the sdf."""
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import CompactBatch, GridTrainer, expand_compact_on_device

    mi, gt, (R, t) = synth.rgbd_batch(3000, num_kf=4, bound=SMALL_BOUND, seed=9, wall_margin=0.3)
    cb = CompactBatch.from_reference(mi, gt, 0.15, pin=False)
    dev = {k: v.cuda() for k, v in cb.tensors().items()}
    mi_d, gt_d = expand_compact_on_device(dev, 0.15, {})
    assert torch.equal(mi_d["sample_frame_ids"].cpu(), mi["sample_frame_ids"])           # bit-exact
    assert torch.equal(gt_d["sdf_valid"].cpu(), gt["sdf_valid"])
    assert torch.equal(gt_d["sdf_signs"].cpu(), gt["sdf_signs"])
    assert torch.equal(mi_d["weights"].cpu(), mi["weights"])
    bad = {k: v.clone() for k, v in gt.items()}
    bad["sdf_valid"][0, 0, 0] = ~bad["sdf_valid"][0, 0, 0]
    with pytest.raises(ValueError):
        CompactBatch.from_reference(mi, bad, 0.15, pin=False)

    losses, params = [], []
    for fmt in ("reference", "compact"):
        net, _, _ = make_pair(num_poses=4)
        for k in range(4):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        net.lock_pose()
        loss = MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                               grad_method="autograd", eik_trunc_dist=None)
        tr = GridTrainer({"epochs": 3, "learning_rate": 1e-3, "grid_training_mode": "joint"}, net, loss, lambda e: (mi, gt),
                         device="cuda")
        tr.pre_epoch(0)
        sink = torch.zeros(3, 4).pin_memory()
        batch = (mi, gt) if fmt == "reference" else cb
        tr.train_host_batches([batch] * 3, loss_sink=sink)
        torch.cuda.synchronize()
        losses.append(sink.clone())
        params.append([f.detach().clone() for f in net.level_tensors()])
    assert rel_err(losses[1], losses[0]) < 1e-5
    for a, b in zip(params[1], params[0]):
        assert rel_err(a, b) < 1e-4


@pytest.mark.parametrize("n_levels,fdim,scale", [(1, 8, 5), (2, 8, 3), (4, 4, 2), (1, 16, 5), (3, 4, 2), (1, 4, 5)])
def test_mapping_step_other_level_channel_shapes(n_levels, fdim, scale):
    """The fused step on every (levels, channels) instantiation: even numbers of 4-channel groups run the
    two-threads-per-point kernel (halves split by level, or by channel group inside one level), odd ones the
    one-thread kernel.  L1 + free space + second-order eikonal against the oracle."""
    from miso_b200.loss import MisoLossMapping
    net, _, o2 = make_pair(n_levels=n_levels, fdim=fdim, scale=scale, base_cell=0.5)
    mi, gt, (R, t) = _batch(5000)
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    assert net.fused_spec() is not None
    L = MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                        grad_method="autograd", eik_trunc_dist=0.1)
    ld = L.compute(net, _to_cuda(mi), _to_cuda(gt))
    sum(v.mean() for v in ld.values()).backward()
    lo = O.mapping_loss(o2, mi, gt, {k: (R[k], t[k]) for k in range(R.shape[0])}, "L1", 1.0, 0.5, 0.1, 0.15,
                        grad_method="autograd", eik_trunc_dist=0.1)
    sum(lo.values()).backward()
    for k in lo:
        assert rel_err(ld[k], lo[k]) < TOL_G, k
    for l in range(n_levels):
        assert rel_err(net.features[l].feature.grad, o2.features[l].grad) < TOL_G, l


def test_unregistered_keyframe_id_is_loud():
    """a sample whose keyframe id has no 'KF<id>' key (or lies outside the table) must not train with
    some other keyframe's pose.  `compute` raises the reference's assertion (grid_net.py:243); the trainer's sync-free
    step returns a NaN loss (kernel poison word) -- and recovers on the next clean batch."""
    from miso_b200.loss import MisoLossMapping
    net, _, _ = make_pair(num_poses=6)
    mi, gt, (R, t) = _batch(3000)
    for k in (0, 1, 3):                                  # keyframe 2 is never registered
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    L = MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                        grad_method="autograd", eik_trunc_dist=None)
    assert (mi["sample_frame_ids"] == 2).any()
    with pytest.raises(AssertionError, match="Key KF2 not found"):
        L.compute(net, _to_cuda(mi), _to_cuda(gt))
    terms = L.step_into_grads(net, _to_cuda(mi), _to_cuda(gt))
    assert torch.isnan(terms).all()
    bad = {k: v.clone() for k, v in mi.items()}
    bad["sample_frame_ids"][bad["sample_frame_ids"] == 2] = 77      # outside the table
    with pytest.raises(AssertionError, match="Key KF77 not found"):
        L.compute(net, _to_cuda(bad), _to_cuda(gt))
    assert torch.isnan(L.step_into_grads(net, _to_cuda(bad), _to_cuda(gt))).all()
    ok = {k: v.clone() for k, v in mi.items()}
    ok["sample_frame_ids"][ok["sample_frame_ids"] == 2] = 3
    assert torch.isfinite(L.step_into_grads(net, _to_cuda(ok), _to_cuda(gt))).all()


def test_locked_keyframes_get_no_pose_gradient_in_mixed_batches():
    """ADVICE r01 / grid_net.py:209-215: unlock_pose(); lock_all_pose_indices(); unlock_pose_index(k) -- the
    reference's track_window pattern -- must send pose gradients to keyframe k only, although the batch holds
    samples of every keyframe; values against the oracle's per-keyframe restatement."""
    from miso_b200.loss import MisoLossMapping
    net, o1, _ = make_pair(num_poses=4)
    mi, gt, (R, t) = _batch(3000)
    for k in range(4):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.lock_feature()
    net.unlock_pose()
    net.lock_all_pose_indices()
    net.unlock_pose_index(2)
    L = MisoLossMapping(loss_type="L2", weight_sdf=1.0, weight_eik=0.0, weight_fs=0.5, trunc_dist=0.15)
    ld = L.compute(net, _to_cuda(mi), _to_cuda(gt))
    sum(ld.values()).backward()
    gr, gtr = net.rotation_corrections.grad, net.translation_corrections.grad
    for k in (0, 1, 3):
        assert torch.count_nonzero(gr[k]) == 0 and torch.count_nonzero(gtr[k]) == 0
    assert torch.count_nonzero(gr[2]) > 0 and torch.count_nonzero(gtr[2]) > 0
    # oracle: keyframe 2's pose through (w, tau) with autograd, the others constant
    w = torch.zeros(1, 3, requires_grad=True)
    tau = torch.zeros(3, 1, requires_grad=True)
    poses = {k: (R[k], t[k]) for k in range(4)}
    poses[2] = O.apply_pose_correction(R[2], t[2], w, tau)
    lo = O.mapping_loss(o1, mi, gt, poses, "L2", 1.0, 0.0, 0.5, 0.15)
    sum(lo.values()).backward()
    assert rel_err(gr[2], w.grad[0]) < TOL_G and rel_err(gtr[2], tau.grad) < TOL_G


@pytest.mark.parametrize("loss_type,eik_trunc,n", [("L1", None, 6000), ("L2", 0.1, 6000), ("L1", None, 50000)])
def test_mapping_step_fused_finite_difference_eikonal(loss_type, eik_trunc, n):
    """grad_method='finitediff' (the shipped default, scannet.yaml:48-49) on the FUSED path (fixed decoder): every
    term and d(total)/d(grid) against the oracle's restatement of diff.py:18-26 + loss.py:754-813."""
    from miso_b200.loss import MisoLossMapping
    from helpers import drop_fragile_points
    net, o1, _ = make_pair()
    mi, gt, (R, t) = _batch(n + n // 4)
    # samples within rounding distance of a ReLU / L1 kink (at x or at one of the six displaced points) get a
    # one-sided derivative that depends on the summation order; one such flip is ~1e-3 of the gradient norm at this
    # batch size, so they are taken out of the batch for BOTH sides (tests/test_gpu_baseline_sizes.py, module doc)
    mi, gt, _ = drop_fragile_points(o1, mi, gt, (R, t), n, 0.15, fd_eps=0.024)
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    L = MisoLossMapping(loss_type=loss_type, weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                        grad_method="finitediff", finite_diff_eps=0.024, eik_trunc_dist=eik_trunc)
    assert L._fused_ok(net)
    ld = L.compute(net, _to_cuda(mi), _to_cuda(gt))
    sum(v.mean() for v in ld.values()).backward()
    lo = O.mapping_loss(o1, mi, gt, {k: (R[k], t[k]) for k in range(R.shape[0])}, loss_type, 1.0, 0.5, 0.1, 0.15,
                        finite_diff_eps=0.024, grad_method="finitediff", eik_trunc_dist=eik_trunc)
    sum(lo.values()).backward()
    assert set(ld) == set(lo)
    for k in lo:
        assert rel_err(ld[k], lo[k]) < TOL_G, k
    for l in range(2):
        assert rel_err(net.features[l].feature.grad, o1.features[l].grad) < TOL_G, l


def test_slab_sharded_fit_single_rank_equals_trainer():
    """miso_b200.sharded_fit on one rank (slab = the whole level): device-side slab selection + device sample count +
    slab Adam reproduce GridTrainer.train_step; and a HALF slab keeps exactly the samples the ownership rule names."""
    from miso_b200.loss import MisoLossMapping
    from miso_b200.sharded_fit import SlabShardedFit, plane_of_points
    from miso_b200.trainer import GridTrainer
    mi, gt, (R, t) = _batch(20000)
    nets = []
    for _ in range(2):
        net, _, _ = make_pair()
        for k in range(R.shape[0]):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        net.lock_pose()
        nets.append(net)
    mk = lambda: MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                                 grad_method="autograd", eik_trunc_dist=0.1)
    tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, nets[0], mk(), None, device="cuda")
    fit = SlabShardedFit(nets[1], mk(), lr=1e-3, rank=0, world=1)
    dmi, dgt = _to_cuda(mi), _to_cuda(gt)
    assert fit.calibrate(dmi) == [0, fit.Z]
    for _ in range(3):
        a = tr.train_step(dmi, dgt)
        b = fit.step(dmi, dgt)
        assert rel_err(b, a) < 1e-5
    assert int(fit._bufs["count"].item()) == 20000
    for pa, pb in zip(nets[0].level_tensors(), nets[1].level_tensors()):
        assert rel_err(pb, pa) < 1e-5
    # ownership rule of a half slab, bit-exact against the torch restatement
    half = SlabShardedFit(nets[1], mk(), lr=1e-3, rank=1, world=2, bounds=[0, fit.Z // 2, fit.Z])
    import torch.distributed as dist
    assert not dist.is_initialized()
    half._exchange_and_update = lambda *a, **k: None          # selection + kernel only: no process group here
    half.step(dmi, dgt)
    assert half.axis == 2 and fit.axis == 2                    # a single rank keeps the native z-slowest layout
    ids = mi["sample_frame_ids"][0, :, 0]
    zw = torch.einsum("nj,nj->n", R[ids][:, 2, :], mi["coords_frame"][0]) + t[ids][:, 2, 0]
    plane = plane_of_points(zw.cuda(), SMALL_BOUND[2][0], SMALL_BOUND[2][1], fit.Z)
    want = torch.nonzero(plane >= fit.Z // 2)[:, 0]
    n = int(half._bufs["count"].item())
    assert abs(n - want.numel()) <= 2          # fma contraction in the torch restatement can move a boundary sample
    got_sdf = torch.sort(half._bufs["sdf"][:n]).values
    if n == want.numel():
        assert torch.equal(got_sdf, torch.sort(gt["sdf"][0, :, 0].cuda()[want]).values)


def test_graphed_train_step_equals_eager_steps():
    """GridTrainer.graphed_train_step (one CUDA-graph launch per step, Adam step counters on the device) reproduces
    the eager launch sequence bit for bit in the losses it returns and to atomics-order noise in the parameters; the
    device-step Adam matches torch.optim.Adam's bias correction (checked against the oracle after 6 steps)."""
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import GridTrainer
    mi, gt, (R, t) = _batch(8000)
    dmi, dgt = _to_cuda(mi), _to_cuda(gt)
    mk = lambda: MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                                 grad_method="autograd", eik_trunc_dist=0.1)
    out = []
    for graph in (False, True):
        net, _, o2 = make_pair()
        for k in range(R.shape[0]):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        net.lock_pose()
        tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint", "cuda_graph": graph}, net, mk(), None,
                         device="cuda")
        step = (lambda: tr.graphed_train_step(dmi, dgt)) if graph else (lambda: tr.train_step(dmi, dgt))
        losses = [step().clone() for _ in range(6)]
        out.append((torch.stack(losses), [p.detach().clone() for p in net.level_tensors()]))
        assert tr.total_steps == 6
    assert rel_err(out[1][0], out[0][0]) < 1e-6
    for a, b in zip(out[1][1], out[0][1]):
        assert rel_err(a, b) < 1e-5
    opt = torch.optim.Adam(list(o2.features.parameters()), lr=1e-3)
    for _ in range(6):
        opt.zero_grad()
        sum(O.mapping_loss(o2, mi, gt, {k: (R[k], t[k]) for k in range(R.shape[0])}, "L1", 1.0, 0.5, 0.1, 0.15,
                           grad_method="autograd", eik_trunc_dist=0.1).values()).backward()
        opt.step()
    for l in range(2):
        assert rel_err(out[1][1][l], o2.features[l]) < TOL_G


def test_graphed_train_step_with_trainable_decoder():
    """decoder.fix: False under the CUDA-graph trainer: the decoder-gradient pass and the decoder's Adam launches are
    captured with the step; six graph launches equal six eager steps (losses, grids, decoder weights)."""
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import GridTrainer
    mi, gt, (R, t) = _batch(5000, seed=13)
    dmi, dgt = _to_cuda(mi), _to_cuda(gt)
    mk = lambda: MisoLossMapping(loss_type="L2", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                                 grad_method="autograd", eik_trunc_dist=None)
    out = []
    for graph in (False, True):
        net, _, _ = make_pair(fix=False)
        for k in range(R.shape[0]):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        net.lock_pose()
        tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint", "cuda_graph": graph}, net, mk(), None,
                         device="cuda")
        step = (lambda: tr.graphed_train_step(dmi, dgt)) if graph else (lambda: tr.train_step(dmi, dgt))
        losses = [step().clone() for _ in range(6)]
        out.append((torch.stack(losses), [p.detach().clone() for p in net.level_tensors()],
                    [p.detach().clone() for p in net.decoder.parameters()]))
    w0 = list(synth.decoder_weights(8, seed=0).values())
    assert rel_err(out[1][0], out[0][0]) < 1e-5
    for a, b in zip(out[1][1], out[0][1]):
        assert rel_err(a, b) < 1e-5
    for a, b, c in zip(out[1][2], out[0][2], w0):
        assert rel_err(a, b) < 1e-5
        if c.numel() > 1:
            assert rel_err(a, c) > 1e-3      # and the decoder was trained


def test_nan_total_skips_the_update_on_the_device():
    """grid_opt/trainer.py:214-217 (`if not isnan(total): backward(); step()`): a step whose total is NaN (here: a
    keyframe without a pose poisons it) leaves parameters, moments and the Adam step counter untouched and clears the
    gradient; the next clean step continues as if the bad one never happened (cuda_graph trainer, device-side gate)."""
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import GridTrainer
    mi, gt, (R, t) = _batch(4000)
    bad = {k: v.clone() for k, v in mi.items()}
    bad["sample_frame_ids"][0, :50, 0] = 99
    mk = lambda: MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                                 grad_method="autograd", eik_trunc_dist=None)
    params = []
    for inject in (False, True):
        net, _, _ = make_pair()
        for k in range(R.shape[0]):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        net.lock_pose()
        tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint", "cuda_graph": True}, net, mk(), None,
                         device="cuda")
        tr.train_step(_to_cuda(mi), _to_cuda(gt))
        if inject:
            before = [p.detach().clone() for p in net.level_tensors()]
            terms = tr.train_step(_to_cuda(bad), _to_cuda(gt))
            assert torch.isnan(terms).all()
            for p, b in zip(net.level_tensors(), before):
                assert torch.equal(p.detach(), b)
                assert torch.count_nonzero(p.grad) == 0
        tr.train_step(_to_cuda(mi), _to_cuda(gt))
        params.append([p.detach().clone() for p in net.level_tensors()])
    for a, b in zip(params[0], params[1]):
        assert rel_err(b, a) < 1e-5


def test_slab_sharded_fit_y_axis_layout_equals_trainer():
    """The slab level stored y-slowest (SlabShardedFit picks z or y, whichever balances the batch): the fused kernels
    take the permuted strides as they are, results equal the channels_last_3d run; a y half-slab owns the samples the
    ownership rule names; restore_layout() returns to channels_last_3d with the same values."""
    from miso_b200.loss import MisoLossMapping
    from miso_b200.sharded_fit import SlabShardedFit, plane_of_points
    from miso_b200.trainer import GridTrainer
    mi, gt, (R, t) = _batch(20000)
    nets = []
    for _ in range(2):
        net, _, _ = make_pair()
        for k in range(R.shape[0]):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        net.lock_pose()
        nets.append(net)
    mk = lambda: MisoLossMapping(loss_type="L2", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.5, trunc_dist=0.15,
                                 grad_method="autograd", eik_trunc_dist=None)
    tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, nets[0], mk(), None, device="cuda")
    fit = SlabShardedFit(nets[1], mk(), lr=1e-3, rank=0, world=1)
    dmi, dgt = _to_cuda(mi), _to_cuda(gt)
    fit.calibrate(dmi, axes=(1,))
    fine = nets[1].level_tensors()[1]
    assert fit.axis == 1 and fine.stride(3) > fine.stride(2) > fine.stride(4) > fine.stride(1) == 1     # y slowest
    for _ in range(3):
        assert rel_err(fit.step(dmi, dgt), tr.train_step(dmi, dgt)) < 1e-5
    for pa, pb in zip(nets[0].level_tensors(), nets[1].level_tensors()):
        assert rel_err(pb, pa) < 1e-5
    Y = fit.Z
    half = SlabShardedFit(nets[1], mk(), lr=1e-3, rank=0, world=2, bounds=[0, Y // 2, Y])
    assert half.axis == 1                                      # picked up from the strides
    half._exchange_and_update = lambda *a, **k: None
    half.step(dmi, dgt)
    ids = mi["sample_frame_ids"][0, :, 0]
    yw = torch.einsum("nj,nj->n", R[ids][:, 1, :], mi["coords_frame"][0]) + t[ids][:, 1, 0]
    want = int((plane_of_points(yw.cuda(), SMALL_BOUND[1][0], SMALL_BOUND[1][1], Y) < Y // 2).sum())
    assert abs(int(half._bufs["count"].item()) - want) <= 2
    vals = fine.detach().clone()
    fit.restore_layout()
    fine2 = nets[1].level_tensors()[1]
    assert fine2.stride(2) > fine2.stride(3) and torch.equal(fine2.detach(), vals)
