"""GPU parity of the batched alignment kernel (C-ABI section 3) against the CPU oracle's restatement
of pairwise_loss_latent (grid_opt/align/miso.py:116-211), check_submap_intersection
(grid_atlas.py:405-420) and generic_align_multiple_submaps (align/base.py:89-163).
Tolerances: loss 1e-5 rel (forward class), pose gradients and Adam pose updates 1e-4 rel,
in-bound masks / valid counts / alignment-sample ordering bit-exact."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from miso_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

BOUND = [[-4.0, 4.0], [-2.0, 2.0], [-4.0, 4.0]]


def build_atlases(num_submaps=3, base_cell=1.0, scale=2, seed=0, perturb=True, spacing=(4.0, 3.0)):
    """(GridAtlas on cuda, OracleAtlas on cpu) with identical grids sampled from one smooth latent field."""
    from miso_b200.models import GridAtlas
    cfg = synth.model_cfg(BOUND, base_cell_size=base_cell, per_level_scale=scale, num_poses=1)
    Rt, tt = synth.submap_layout(num_submaps, spacing=spacing)
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=4.0, trans_m=0.3) if perturb else (Rt, tt)
    atlas = GridAtlas(cfg, device="cuda")
    shapes = O.level_shapes(BOUND, base_cell, scale, 2, 4)
    subs = []
    for i in range(num_submaps):
        atlas.add_submap(torch.tensor(BOUND), Rp[i], tp[i])
        feats = synth.fill_submap_from_field(shapes, BOUND, Rt[i], tt[i])
        # zero a slab so precompute_coordinates_for_alignment has something to prune
        feats[0][:, :, :, :, :1] = 0
        feats[1][:, :, :, :, :2] = 0
        sm = atlas.get_submap(i)
        sm.decoder.load_state_dict(synth.decoder_weights(8))
        for l in range(2):
            with torch.no_grad():
                sm.features[l].feature.copy_(feats[l].cuda())
        subs.append(O.OracleGridNet(BOUND, feats, None))
    oat = O.OracleAtlas(subs, Rp, tp)
    return atlas, oat


def test_precompute_coordinates_bit_exact():
    atlas, oat = build_atlases(2)
    atlas.precompute_coordinates_for_alignment()
    oat.precompute([0, 1])
    for i in range(2):
        for l in range(2):
            a = atlas.coordinates_for_alignment(i, l).cpu()
            b = oat.coords[(i, l)]
            assert a.shape == b.shape and a.shape[0] > 0
            assert torch.equal(a, b)


@pytest.mark.parametrize("level", [0, 1])
def test_pairwise_loss_and_pose_grads(level):
    from miso_b200.align import pairwise_loss_latent
    atlas, oat = build_atlases(3)
    atlas.precompute_coordinates_for_alignment()
    oat.precompute([0, 1])
    for (s, d) in [(0, 1), (1, 2), (0, 2)]:
        for p in list(atlas.rotation_corrections) + list(atlas.translation_corrections):
            p.grad = None
        ld = pairwise_loss_latent(atlas, None, s, d, level=level, device="cuda")
        (key, val), = ld.items()
        assert key == f"align_latent_level{level}_{s}_{d}"
        Rs, ts = oat.updated_submap_pose(s)
        Rd, td = oat.updated_submap_pose(d)
        for q in oat.rot + oat.tra:
            q.grad = None
        lo = O.pairwise_loss_latent(oat.submaps[s], oat.submaps[d], oat.coords[(s, level)], Rs, ts, Rd, td, level)
        assert rel_err(val, lo) < 1e-5
        if lo.requires_grad:
            val.backward()
            lo.backward()
            for i in (s, d):
                assert rel_err(atlas.rotation_corrections[i].grad, oat.rot[i].grad) < 1e-4
                assert rel_err(atlas.translation_corrections[i].grad, oat.tra[i].grad) < 1e-4


def test_mask_count_intersection_bit_exact():
    from miso_b200.align import AlignBatch
    atlas, oat = build_atlases(3)
    atlas.precompute_coordinates_for_alignment()
    oat.precompute([1])
    pairs = [(0, 1), (0, 2), (1, 2)]
    batch = AlignBatch(atlas, pairs, level=1, want_masks=True, check_intersection=True)
    poses = batch.pair_poses()
    batch.update_intersections(poses)
    out = batch.launch(poses)
    for i, (s, d) in enumerate(pairs):
        Rs, ts = oat.updated_submap_pose(s)
        Rd, td = oat.updated_submap_pose(d)
        _, mask, idx = O.pairwise_loss_latent(oat.submaps[s], oat.submaps[d], oat.coords[(s, 1)], Rs, ts, Rd, td, 1,
                                              return_aux=True)
        inter = bool(O.check_submap_intersection(oat.submaps[s], oat.submaps[d], Rs, ts, Rd, td))
        assert bool(batch.enabled[i].item()) == inter
        if inter:
            assert torch.equal(batch.masks[i].cpu().bool(), mask[:, 0])
            assert int(out[i, 1].item()) == int(mask.sum())
            # nonzero() of the mask reproduces valid_indices (miso.py:184) in order
            assert torch.equal(torch.nonzero(batch.masks[i], as_tuple=False)[:, 0].cpu(), idx if idx is not None else torch.empty(0, dtype=torch.long))


def test_no_valid_points_gives_zero():
    from miso_b200.align import AlignBatch
    atlas, _ = build_atlases(2, spacing=(100.0, 100.0), perturb=False)
    atlas.precompute_coordinates_for_alignment()
    batch = AlignBatch(atlas, [(0, 1)], level=0, check_intersection=False)
    loss = batch.losses(3000.0)
    assert float(loss[0]) == 0.0
    loss.sum().backward()  # gradients exist and are zero
    assert torch.count_nonzero(atlas.rotation_corrections[1].grad) == 0


def test_align_iterations_match_oracle_adam():
    """k iterations of generic_align_multiple_submaps: total loss per iteration and the resulting pose
    corrections (the Adam update) against the oracle loop."""
    from miso_b200.align import generic_align_multiple_submaps
    atlas, oat = build_atlases(3)
    atlas.precompute_coordinates_for_alignment()
    oat.precompute([0])
    info = generic_align_multiple_submaps(atlas, None, ("latent", None), num_iters=4, lr=1e-2, level=0)
    hist = O.align_multiple_submaps(oat, level=0, num_iters=4, lr=1e-2)
    got = info["losses"].tolist()
    assert len(got) == len(hist) == 5
    assert np.allclose(got, hist, rtol=1e-4), (got, hist)
    for i in range(1, 3):
        assert rel_err(atlas.rotation_corrections[i], oat.rot[i]) < 1e-4
        assert rel_err(atlas.translation_corrections[i], oat.tra[i]) < 1e-4
    assert torch.count_nonzero(atlas.rotation_corrections[0]) == 0  # submap 0 stays fixed


def test_hierarchical_alignment_decreases_loss():
    """align_multiple_submaps_hierarchical (miso.py:217-322) over levels [0, 1]: the summed pair loss must
    drop at each level and the poses stay finite (pose-error recovery itself depends on the data; parity of
    every iterate with the oracle is asserted in test_align_iterations_match_oracle_adam)."""
    from miso_b200.align import align_multiple_submaps_hierarchical
    atlas, _ = build_atlases(3)
    info = align_multiple_submaps_hierarchical(atlas, None, level_iters=40, lr=1e-2, latent_levels=[0, 1],
                                               skip_finetune=True)
    for lvl in (0, 1):
        h = info[f"hier_latent_level{lvl}_L2"]["losses"]
        assert len(h) == 41                       # num_iters + 1 iterations (base.py:127)
        assert torch.isfinite(h).all()
        if lvl == 0:
            assert float(h[-1]) < 0.9 * float(h[0])
    for i in range(3):
        R, t = atlas.updated_submap_pose(i)
        assert torch.isfinite(R).all() and torch.isfinite(t).all()


def test_gauss_newton_normal_equations():
    """J^T J / J^T r of the latent residual wrt a right-multiplied dst twist, against autograd."""
    from miso_b200.align import AlignBatch, gauss_newton_dst_step
    atlas, oat = build_atlases(2)
    atlas.precompute_coordinates_for_alignment()
    oat.precompute([0])
    batch = AlignBatch(atlas, [(0, 1)], level=0, check_intersection=False)
    delta, H, g = gauss_newton_dst_step(batch, lm_lambda=0.0)
    # oracle: residual vector as a function of a twist xi = (w, tau) applied as R_d Exp(w), t_d + tau
    Rs, ts = oat.updated_submap_pose(0)
    Rd0, td0 = oat.updated_submap_pose(1)
    Rs, ts, Rd0, td0 = Rs.detach().double(), ts.detach().double(), Rd0.detach().double(), td0.detach().double()
    src = O.OracleGridNet(BOUND, [f.double() for f in oat.submaps[0].features], None, second_order=True)
    dst = O.OracleGridNet(BOUND, [f.double() for f in oat.submaps[1].features], None, second_order=True)
    p = oat.coords[(0, 0)].double()

    def resid(xi):
        Rd = Rd0 @ torch.linalg.matrix_exp(O.hat(xi[None, :3])[0])
        td = td0 + xi[3:, None]
        q = O.transfrom_points_from(O.transform_points_to(p, Rs, ts), Rd, td)
        m = O.coords_in_bound(q.detach(), dst.bound.double())[:, 0]
        r = src.query_feature(p[m])[:, :4] - dst.query_feature(q[m])[:, :4]
        return r.reshape(-1)

    xi0 = torch.zeros(6, dtype=torch.float64)
    J = torch.autograd.functional.jacobian(resid, xi0)
    r0 = resid(xi0)
    assert rel_err(H[0], J.T @ J) < 1e-4
    assert rel_err(g[0], J.T @ r0) < 1e-4


def test_cuda_graph_replay_matches_eager():
    """One whole alignment iteration captured as a CUDA graph must reproduce the eager iterates."""
    from miso_b200.align import generic_align_multiple_submaps
    a1, _ = build_atlases(3)
    a2, _ = build_atlases(3)
    for a in (a1, a2):
        a.precompute_coordinates_for_alignment()
    i1 = generic_align_multiple_submaps(a1, None, ("latent", None), num_iters=8, lr=1e-2, level=0)
    i2 = generic_align_multiple_submaps(a2, None, ("latent", None), num_iters=8, lr=1e-2, level=0, use_cuda_graph=True)
    assert len(i1["losses"]) == len(i2["losses"]) == 9
    assert np.allclose(i1["losses"].numpy(), i2["losses"].numpy(), rtol=1e-4)
    for i in range(1, 3):
        assert rel_err(a2.rotation_corrections[i], a1.rotation_corrections[i]) < 1e-4
        assert rel_err(a2.translation_corrections[i], a1.translation_corrections[i]) < 1e-4


@pytest.mark.parametrize("level,graph", [(0, False), (1, False), (0, True)])
def test_fused_pose_glue_matches_torch_glue(level, graph):
    """The five-launch iteration (compose / intersect / align / pose-gradient / Adam kernels, csrc/poseopt.cu)
    against the same loop with the reference's torch glue (so3_exp_map autograd + torch.optim.Adam): per-iteration
    losses and the pose corrections after 12 iterations, starting from non-zero corrections on one submap so the
    un-clamped branch of so3_exp_map's derivative is exercised too."""
    from miso_b200.align import generic_align_multiple_submaps
    atlases = []
    for _ in range(2):
        a, _o = build_atlases(3)
        a.precompute_coordinates_for_alignment()
        with torch.no_grad():
            a.rotation_corrections[2].copy_(torch.tensor([[0.03, -0.02, 0.015]]))
            a.translation_corrections[2].copy_(torch.tensor([[0.05], [-0.02], [0.01]]))
        atlases.append(a)
    i_t = generic_align_multiple_submaps(atlases[0], None, ("latent", None), num_iters=11, lr=1e-2, level=level,
                                         fused_pose_glue=False)
    i_f = generic_align_multiple_submaps(atlases[1], None, ("latent", None), num_iters=11, lr=1e-2, level=level,
                                         fused_pose_glue=True, use_cuda_graph=graph)
    assert len(i_t["losses"]) == len(i_f["losses"]) == 12
    assert np.allclose(i_f["losses"].numpy(), i_t["losses"].numpy(), rtol=1e-4), (i_f["losses"], i_t["losses"])
    for i in range(1, 3):
        assert rel_err(atlases[1].rotation_corrections[i], atlases[0].rotation_corrections[i]) < 1e-4
        assert rel_err(atlases[1].translation_corrections[i], atlases[0].translation_corrections[i]) < 1e-4
    assert torch.count_nonzero(atlases[1].rotation_corrections[0]) == 0
    assert torch.count_nonzero(atlases[1].translation_corrections[0]) == 0


def test_intersection_counts_lattice_path_equals_per_vertex_path():
    """The lattice-row path of miso_align_intersections (whole row segments decided by their end points, margin two
    orders above the rounding error) must give EXACTLY the per-vertex counts, for overlapping, barely touching and
    disjoint pairs, and agree with torch's count up to vertices that sit on the bound within rounding."""
    from miso_b200 import geometry as G
    from miso_b200.align import AlignBatch
    atlas, _ = build_atlases(4, base_cell=0.5, scale=4, spacing=(5.0, 3.5))
    atlas.precompute_coordinates_for_alignment()
    with torch.no_grad():
        atlas.rotation_corrections[3].copy_(torch.tensor([[0.4, -0.3, 0.2]]))      # a strongly rotated submap
        atlas.translation_corrections[2].copy_(torch.tensor([[30.0], [0.0], [0.0]]))  # far away: disjoint from the rest
        atlas.translation_corrections[1].copy_(torch.tensor([[2.9], [0.0], [0.0]]))   # a thin sliver of overlap with 0
    pairs = [(s, d) for s in range(4) for d in range(4) if s != d]
    counts = []
    for lattice in (True, False):
        b = AlignBatch(atlas, pairs, level=0, check_intersection=True, lattice_rows=lattice)
        b.update_intersections(b.pair_poses())
        torch.cuda.synchronize()
        counts.append(b.counts[:len(pairs)].clone())
    assert torch.equal(counts[0], counts[1]), (counts[0], counts[1])
    assert int((counts[0] > 0).sum()) >= 4 and int((counts[0] == 0).sum()) >= 1
    for i, (s, d) in enumerate(pairs):
        v = atlas.get_submap(s).features[-1].vertex_positions().cuda()
        Rs, ts = atlas.updated_submap_pose(s)
        Rd, td = atlas.updated_submap_pose(d)
        q = G.transfrom_points_from(G.transform_points_to(v, Rs, ts), Rd, td)
        ref = int(G.coords_in_bound(q, atlas.get_submap(d).bound.cuda()).sum())
        assert abs(int(counts[0][i]) - ref) <= max(4, ref // 20000), (s, d, int(counts[0][i]), ref)


def test_pose_adam_skips_submaps_without_a_gradient():
    """A submap none of whose pairs gave a gradient this iteration has .grad None in the reference's loop (its loss does
    not depend on the submap), so torch.optim.Adam leaves it alone: no moment decay, no step increment, no momentum
    drift.  miso_align_pose_adam with `contrib` / `submap_steps` against torch.optim.Adam with per-iteration None
    gradients (submap 2 contributes in iterations 0 and 2 only, submap 3 never; submap 0 is fixed)."""
    from miso_b200 import _lib
    lib = _lib.load()
    S, iters = 4, 4
    g = torch.Generator().manual_seed(3)
    w0 = [torch.randn(1, 3, generator=g) * 0.1 for _ in range(S)]
    t0 = [torch.randn(3, 1, generator=g) * 0.1 for _ in range(S)]
    grads = torch.randn(iters, S, 6, generator=g)
    contrib = torch.tensor([[2, 3, 1, 0], [2, 2, 0, 0], [1, 1, 1, 0], [1, 1, 0, 0]], dtype=torch.float32)
    # reference semantics
    pw = [torch.nn.Parameter(x.clone()) for x in w0]
    pt = [torch.nn.Parameter(x.clone()) for x in t0]
    opt = torch.optim.Adam(pw[1:] + pt[1:], lr=1e-2)
    for it in range(iters):
        for s in range(1, S):
            has = contrib[it, s] > 0
            pw[s].grad = grads[it, s, :3].reshape(1, 3).clone() if has else None
            pt[s].grad = grads[it, s, 3:].reshape(3, 1).clone() if has else None
        opt.step()
    # kernel
    dw = [x.clone().cuda() for x in w0]
    dt = [x.clone().cuda() for x in t0]
    w_ptrs = torch.tensor([x.data_ptr() for x in dw], dtype=torch.int64, device="cuda")
    t_ptrs = torch.tensor([x.data_ptr() for x in dt], dtype=torch.int64, device="cuda")
    m = torch.zeros(S, 6, device="cuda")
    v = torch.zeros(S, 6, device="cuda")
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    steps = torch.zeros(S, dtype=torch.int32, device="cuda")
    for it in range(iters):
        gi, ci = grads[it].contiguous().cuda(), contrib[it].contiguous().cuda()
        _lib.check(lib.miso_align_pose_adam(w_ptrs.data_ptr(), t_ptrs.data_ptr(), S, gi.data_ptr(), m.data_ptr(), v.data_ptr(),
                                            counter.data_ptr(), 1e-2, 0.9, 0.999, 1e-8, ci.data_ptr(), steps.data_ptr(),
                                            _lib.stream_ptr(torch.device("cuda", 0))), "align_pose_adam")
    torch.cuda.synchronize()
    assert int(counter.item()) == iters and steps.tolist() == [0, 4, 2, 0]
    for s in range(S):
        assert rel_err(dw[s], pw[s]) < 1e-6 and rel_err(dt[s], pt[s]) < 1e-6, s
    assert torch.equal(dw[3].cpu(), w0[3]) and torch.equal(dw[0].cpu(), w0[0])
