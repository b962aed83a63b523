"""CPU: host-side logic of the product package that needs no GPU -- geometry glue, descriptors,
synthetic generators, error behaviour (no CPU fallback), partitioning helpers."""
import numpy as np
import pytest
import torch

from miso_b200 import dist as mdist
from miso_b200 import field, geometry, synth
from oracle import oracle as O


def test_geometry_matches_oracle_restatement():
    w = torch.randn(20, 3) * 0.5
    assert torch.allclose(geometry.so3_exp_map(w), O.so3_exp_map(w), atol=1e-7)
    assert torch.allclose(geometry.hat(w), O.hat(w))
    R = geometry.so3_exp_map(torch.randn(1, 3))[0]
    t = torch.randn(3, 1)
    p = torch.randn(50, 3)
    assert torch.equal(geometry.transform_points_to(p, R, t), O.transform_points_to(p, R, t))
    assert torch.equal(geometry.transfrom_points_from(p, R, t), O.transfrom_points_from(p, R, t))
    back = geometry.transfrom_points_from(geometry.transform_points_to(p, R, t), R, t)
    assert torch.allclose(back, p, atol=1e-5)
    b = torch.tensor([[-1.0, 1.0], [-2.0, 2.0], [0.0, 3.0]])
    pts = torch.tensor([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0001], [-1.0, -2.0, 0.0], [0.0, 0.0, -1e-6]])
    assert geometry.coords_in_bound(pts, b)[:, 0].tolist() == [True, False, True, False]   # inclusive bounds
    Rn, tn = geometry.apply_pose_correction(R, t, torch.zeros(1, 3), torch.zeros(3, 1))
    assert torch.allclose(Rn, R, atol=1e-7) and torch.equal(tn, t)


def test_field_descriptor_from_channels_last_tensor():
    f = torch.randn(1, 4, 5, 6, 7).contiguous(memory_format=torch.channels_last_3d)
    fld = field.make_field([f], [-1, 1, -2, 2, -3, 3], None, ignore_mask=0)
    lv = fld.level[0]
    assert (lv.X, lv.Y, lv.Z, lv.C) == (7, 6, 5, 4)
    assert (lv.sC, lv.sX, lv.sY, lv.sZ) == (1, 4, 28, 168)
    assert list(fld.bound) == [-1, 1, -2, 2, -3, 3]
    p = torch.nn.Parameter(torch.randn(1, 4, 3, 3, 3))
    field.to_channels_last_3d_(p)
    assert p.stride()[1] == 1 and p.shape == (1, 4, 3, 3, 3)
    with pytest.raises(RuntimeError):
        field.make_field([torch.randn(4, 5, 6, 7)], [0, 1, 0, 1, 0, 1])
    with pytest.raises(RuntimeError):
        field.make_field([f] * 5, [0, 1, 0, 1, 0, 1])


def test_no_cpu_fallback():
    """The product path must fail loudly off-GPU instead of silently computing on the CPU."""
    from miso_b200 import cuda_gridsample as cu
    with pytest.raises(RuntimeError):
        cu.grid_sample_3d(torch.randn(1, 4, 3, 3, 3), torch.zeros(1, 2, 1, 1, 3), padding_mode="zeros", align_corners=False)
    with pytest.raises(RuntimeError):
        field.field_features_raw([torch.randn(1, 4, 3, 3, 3)], [0, 1, 0, 1, 0, 1], torch.zeros(5, 3))
    with pytest.raises(NotImplementedError):
        cu.grid_sample_2d(torch.randn(1, 1, 3, 3), torch.zeros(1, 1, 1, 2))


def test_package_never_imports_oracle():
    import os, re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "miso_b200")
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            src = open(os.path.join(root, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_model_mirror_shapes_and_state_dict_keys():
    from miso_b200.models import GridNet
    cfg = synth.model_cfg(synth.SCANNET_SUBMAP_BOUND, num_poses=3)
    net = GridNet(cfg, device="cpu")
    assert tuple(net.features[0].feature.shape) == (1, 4, 40, 20, 40)
    assert tuple(net.features[1].feature.shape) == (1, 4, 200, 100, 200)
    assert tuple(net.feature_stability[1].feature.shape) == (1, 1, 200, 100, 200)
    assert net.features[1].feature.stride()[1] == 1                          # channels_last_3d storage
    keys = set(net.state_dict().keys())
    for k in ["features.0.feature", "features.1.feature", "feature_stability.0.feature", "decoder.network.0.weight",
              "decoder.network.4.bias", "rotation_corrections", "translation_corrections", "Rwk", "twk"]:
        assert k in keys
    assert sum(p.numel() for p in net.decoder.parameters()) == 4801
    assert [p.shape for p in net.params_at_level(0)] == [net.features[0].feature.shape, net.feature_stability[0].feature.shape]
    assert len(net.params_at_level(2)) == 4
    net.set_initial_kf_pose(1, torch.eye(3), torch.ones(3, 1), kf_key="KF7")
    R, t = net.updated_kf_pose_from_key("KF7")
    assert torch.allclose(R, torch.eye(3)) and torch.equal(t, torch.ones(3, 1))
    with pytest.raises(AssertionError):
        net.set_initial_kf_pose(5, torch.eye(3), torch.ones(3, 1))
    # the level geometry agrees with the oracle's restatement of grid_modules.py:47-57
    assert O.level_shapes(synth.SCANNET_SUBMAP_BOUND, 0.5, 5, 2, 4) == [tuple(f.feature.shape) for f in net.features]
    assert O.level_shapes(synth.NCD_QUAD_BOUND, 1.0, 5, 2, 4)[1] == (1, 4, 100, 450, 450)


def test_vertex_positions_match_oracle_bit_exact():
    from miso_b200.models import GridNet
    b = [[-2.0, 2.0], [-1.0, 1.0], [-2.0, 2.0]]
    net = GridNet(synth.model_cfg(b, num_poses=1), device="cpu")
    for l in range(2):
        assert torch.equal(net.features[l].vertex_positions(), O.vertex_positions(net.features[l].feature.shape, b))


def test_atlas_integer_indexing():
    from miso_b200.models import GridAtlas
    b = [[-2.0, 2.0], [-1.0, 1.0], [-2.0, 2.0]]
    atlas = GridAtlas(synth.model_cfg(b, num_poses=4), device="cpu")
    for s in range(3):
        atlas.add_submap(torch.tensor(b), torch.eye(3), torch.zeros(3, 1), num_poses=4)
        for k in range(2 + s):
            atlas.add_kf(torch.eye(3), torch.zeros(3, 1))
    assert atlas._kf_id_to_submap_id == [0, 0, 1, 1, 1, 2, 2, 2, 2]
    ids = torch.tensor([8, 0, 3, 5, 2])
    assert atlas.submap_id_for_kf_batch(ids).tolist() == [2, 0, 1, 2, 1]
    assert [atlas.anchor_kf_for_submap(s) for s in range(3)] == [0, 2, 5]
    assert atlas.num_keyframes == 9 and atlas.num_submaps == 3
    atlas.set_submap_pose_correction(1, torch.ones(1, 3) * 0.1, torch.ones(3, 1))
    atlas.set_submap_pose(1, torch.eye(3), torch.ones(3, 1))
    assert torch.count_nonzero(atlas.rotation_corrections[1]) == 0    # reset like grid_atlas.py:171-187


def test_synth_generators_are_seeded_and_shaped():
    mi, gt, (R, t) = synth.rgbd_batch(5000, num_kf=5, seed=3)
    mi2, gt2, _ = synth.rgbd_batch(5000, num_kf=5, seed=3)
    assert all(torch.equal(mi[k], mi2[k]) for k in mi) and all(torch.equal(gt[k], gt2[k]) for k in gt)
    assert mi["coords_frame"].shape == (1, 5000, 3) and mi["sample_frame_ids"].dtype == torch.int64
    assert gt["sdf_valid"].dtype == torch.bool and set(gt["sdf_signs"].unique().tolist()) <= {-1.0, 0.0, 1.0}
    # samples transformed by their keyframe pose fall inside the submap bound
    ids = mi["sample_frame_ids"][0, :, 0]
    xw = torch.einsum("nij,nj->ni", R[ids], mi["coords_frame"][0]) + t[ids, :, 0]
    b = torch.tensor(synth.SCANNET_SUBMAP_BOUND)
    assert geometry.coords_in_bound(xw, b).float().mean() > 0.99
    mi, gt, _ = synth.lidar_batch(4000, seed=1)
    assert mi["weights"].min() > 0.5 and gt["sdf"].shape == (1, 4000, 1)


def test_partition_helpers():
    assert mdist.submaps_for_rank(16, 3, 8) == [3, 11]
    assert sorted(sum([mdist.submaps_for_rank(16, r, 8) for r in range(8)], [])) == list(range(16))
    chunks = [mdist.shard_points(1000003, r, 8) for r in range(8)]
    assert chunks[0][0] == 0 and chunks[-1][1] == 1000003
    assert all(chunks[i][1] == chunks[i + 1][0] for i in range(7))
    assert max(e - b for b, e in chunks) - min(e - b for b, e in chunks) <= 1
    pairs = [(s, d) for s in range(16) for d in range(s + 1, 16)]
    owned = [[p for i, p in enumerate(pairs) if mdist.pair_filter(r, 8)(i, p)] for r in range(8)]
    assert sorted(sum(owned, [])) == pairs and len(pairs) == 120
    assert mdist.sharded_loss_scale(250, 1000) == 0.25


def test_balanced_pair_owner_is_deterministic_and_balanced():
    """Cost-balanced alignment-pair ownership: disjoint cover, deterministic, max load <= LPT bound."""
    from miso_b200 import dist as mdist
    rng = np.random.RandomState(0)
    costs = [float(c) for c in rng.randint(0, 4_000_000, size=120) * (rng.rand(120) < 0.35)]
    for w in (1, 2, 4, 8):
        owner = mdist.balanced_pair_owner(costs, w)
        assert owner == mdist.balanced_pair_owner(costs, w)
        assert set(owner) <= set(range(w)) and len(owner) == 120
        loads = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(w)]
        assert max(loads) <= sum(costs) / w + max(costs) + 1e-6          # LPT guarantee
        mine = [[i for i in range(120) if mdist.balanced_pair_filter(costs, r, w)(i, None)] for r in range(w)]
        assert sorted(sum(mine, [])) == list(range(120))


def test_compact_batch_host_side():
    """CompactBatch.from_reference (host side, no GPU): 18 B/point for RGB-D batches (unit weights dropped), weights
    kept for LiDAR batches, refusal when the masks are not the dataset's functions of the sdf or ids overflow int16."""
    from miso_b200.trainer import CompactBatch
    mi, gt, _ = synth.rgbd_batch(3000, num_kf=5, seed=2)
    cb = CompactBatch.from_reference(mi, gt, 0.15, pin=False)
    t = cb.tensors()
    assert set(t) == {"coords", "ids16", "sdf"} and t["ids16"].dtype == torch.int16
    assert sum(v.numel() * v.element_size() for v in t.values()) == 18 * 3000
    assert torch.equal(t["ids16"].long(), mi["sample_frame_ids"][0, :, 0])
    mil, gtl, _ = synth.lidar_batch(4000, num_kf=3, seed=4)
    cbl = CompactBatch.from_reference(mil, gtl, 0.5, pin=False)
    assert "weights" in cbl.tensors() and torch.equal(cbl.tensors()["weights"], mil["weights"][0, :, 0])
    with pytest.raises(ValueError):
        CompactBatch.from_reference(mi, gt, 0.2, pin=False)          # masks were built with trunc 0.15
    big = {k: v.clone() for k, v in mi.items()}
    big["sample_frame_ids"][0, 0, 0] = 40000
    with pytest.raises(ValueError):
        CompactBatch.from_reference(big, gt, 0.15, pin=False)
