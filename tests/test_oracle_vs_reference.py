"""CPU, build container only: re-runs the UNMODIFIED reference live next to the oracle on fresh seeds
(skipped where /root/reference is absent, e.g. on the GPU box; the committed fixtures cover that case)."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from miso_b200 import synth
from oracle import oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")

BOUND = [[-2.0, 2.0], [-1.0, 1.0], [-2.0, 2.0]]


@pytest.fixture(autouse=True)
def _reference_on_path():
    ref_loader.load_reference()


def _ref_net(seed):
    ref_loader.load_reference()
    from grid_opt.models.grid_net import GridNet
    net = GridNet(ref_loader.reference_model_cfg(BOUND, num_poses=4), device="cpu")
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for lvl in net.features:
            lvl.feature.copy_(torch.randn(lvl.feature.shape, generator=g) * 0.1)
    net.decoder.load_state_dict(synth.decoder_weights(8, seed=seed))
    return net


def _oracle_of(net, second_order=False):
    dec = O.make_decoder(8)
    dec.load_state_dict({k.replace("network.", ""): v for k, v in net.decoder.state_dict().items()})
    return O.OracleGridNet(BOUND, [f.feature.data for f in net.features], dec, second_order=second_order)


@pytest.mark.parametrize("seed", [1, 2])
def test_forward_loss_and_grads_equal_reference(seed):
    import grid_opt.loss as rloss
    net = _ref_net(seed)
    om = _oracle_of(net)
    mi, gt, (R, t) = synth.rgbd_batch(3000, num_kf=4, bound=BOUND, seed=seed, wall_margin=0.3)
    for k in range(4):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    L = rloss.MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.0, weight_fs=0.1, trunc_dist=0.15)
    ld = L.compute(net, mi, gt)
    sum(v.mean() for v in ld.values()).backward()
    lo = O.mapping_loss(om, mi, gt, {k: (R[k], t[k]) for k in range(4)}, "L1", 1.0, 0.0, 0.1, 0.15)
    sum(lo.values()).backward()
    for k in lo:
        assert float(ld[k]) == float(lo[k])
    for l in range(2):
        assert torch.equal(net.features[l].feature.grad, om.features[l].grad)


def test_eikonal_fd_equal_and_autograd_first_order_close():
    import grid_opt.loss as rloss
    import grid_opt.diff as rdiff
    net = _ref_net(3)
    om, om2 = _oracle_of(net), _oracle_of(net, second_order=True)
    x = (torch.rand(500, 3) - 0.5) * torch.tensor([3.6, 1.8, 3.6])
    e_ref = rloss.miso_loss_eikonal(net, x, torch.zeros(500, 1), None, "finitediff", 0.024)
    e_or = O.miso_loss_eikonal(om, x, torch.zeros(500, 1), None, "finitediff", 0.024)
    assert float(e_ref) == float(e_or)
    g_ref = rdiff.gradient3d(x.clone().requires_grad_(True), net, "autograd", create_graph=False)
    g_or = O.gradient3d(x.clone().requires_grad_(True), om2, "autograd", create_graph=False)
    assert rel_err(g_or, g_ref) < 1e-6


def test_tracker_normal_equations_follow_reference_formula():
    """Tracker.lm_step (tracker.py:172-197) needs the dataset plumbing, so its formula is exercised on the
    reference model directly with the reference's own helpers."""
    ref_loader.load_reference()
    import grid_opt.diff as rdiff
    import grid_opt.utils.utils_geometry as rgeo
    net = _ref_net(4)
    om2 = _oracle_of(net, second_order=True)
    g = torch.Generator().manual_seed(0)
    xf = (torch.rand(300, 3, generator=g) - 0.5) * torch.tensor([2.0, 1.0, 2.0])
    gt_sdf = torch.randn(300, 1, generator=g) * 0.05
    Rwf = O.so3_exp_map(torch.tensor([[0.1, -0.2, 0.05]]))[0]
    twf = torch.tensor([[0.1], [0.05], [-0.1]])
    xw = rgeo.transform_points_to(xf, Rwf, twf)
    x = xw.clone().requires_grad_(True)
    gw = rdiff.gradient3d(x, net, method="autograd", create_graph=False).detach()
    Rxi = rgeo.transform_points_to(xf, Rwf, torch.zeros_like(twf))
    cT = torch.bmm(ref_loader.hat(Rxi), gw.unsqueeze(-1)).squeeze(-1)
    J = torch.cat((cT @ Rwf, gw), dim=1)
    r = (net(xw) - gt_sdf).detach()
    w = 0.1 / (0.1 + r ** 2) ** 2
    H = J.T @ (w * J) + 1e-4 * torch.eye(6)
    b = J.T @ (w * r)
    Ho, bo, _ = O.lm_normal_equations(om2, xf, gt_sdf, Rwf, twf, loss_type="GM", gm_scale=0.1, lm_lambda=1e-4)
    assert rel_err(Ho, H) < 1e-5 and rel_err(bo, b) < 1e-5


def test_adapter_patches_reference_classes_and_leaves_cpu_behaviour_alone():
    """miso_b200.adapter.patch_reference(): the reference's own GridNet / FeatureGrid / loss gain the CUDA routes, CPU
    tensors keep flowing through the reference's original methods (bit-identical results), unpatch restores them."""
    import grid_opt.diff as rdiff
    import grid_opt.loss as rloss
    import grid_opt.models.grid_modules as rgm
    import grid_opt.models.grid_net as rgn
    from miso_b200 import adapter
    net = _ref_net(3)
    mi, gt, (R, t) = synth.rgbd_batch(2000, num_kf=4, bound=BOUND, seed=3, wall_margin=0.3)
    for k in range(4):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    x = torch.rand(500, 3) * 2 - 1
    L = rloss.MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.0, weight_fs=0.1, trunc_dist=0.15)
    before = (net(x).detach().clone(), net.query_feature(x).detach().clone(),
              {k: float(v) for k, v in L.compute(net, mi, gt).items()})
    originals = (rgn.GridNet.forward, rgm.FeatureGrid.interpolate, rloss.MisoLossMappingBase.compute, rdiff.gradient3d)
    names = adapter.patch_reference()
    try:
        assert {"forward", "interpolate", "query_feature", "compute", "gradient3d"} <= set(names)
        assert rgn.GridNet.forward is not originals[0] and hasattr(rgn.GridNet, "forward_with_gradient")
        assert net.fused_spec() is None                      # CPU model: the fused path does not apply ...
        after = (net(x).detach(), net.query_feature(x).detach(), {k: float(v) for k, v in L.compute(net, mi, gt).items()})
        assert torch.equal(before[0], after[0]) and torch.equal(before[1], after[1]) and before[2] == after[2]   # ... and nothing changes
        Rk, tk = net.all_kf_poses()
        R0, t0 = net.updated_kf_pose(2)
        assert torch.allclose(Rk[2], R0) and torch.equal(tk[2], t0)
    finally:
        adapter.unpatch_reference()
    assert (rgn.GridNet.forward, rgm.FeatureGrid.interpolate, rloss.MisoLossMappingBase.compute, rdiff.gradient3d) == originals
    assert not hasattr(rgn.GridNet, "forward_with_gradient")
