"""CPU: the N>1 partitioning paths with world_size-2 gloo.  The oracle stands in
for the kernels (tests may use it as a checker); what is under test is the host logic in miso_b200.dist:
  * point-sharded fit: per-rank loss/gradients computed with the GLOBAL denominators sum to the
    single-process result after one all_reduce,
  * pair-sharded alignment: round-robin pairs + all_reduce of pose gradients == all pairs on one rank,
  * submap-per-rank: disjoint cover, gather to rank 0 in id order, no collective in between."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from miso_b200 import dist as mdist
from miso_b200 import synth
from oracle import oracle as O

BOUND = [[-2.0, 2.0], [-1.0, 1.0], [-2.0, 2.0]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model(second_order=False):
    shapes = O.level_shapes(BOUND, 0.5, 5, 2, 4)
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(s, generator=g) * 0.1 for s in shapes]
    dec = O.make_decoder(8)
    dec.load_state_dict({k.replace("network.", ""): v for k, v in synth.decoder_weights(8).items()})
    return O.OracleGridNet(BOUND, feats, dec, second_order=second_order)


def _sharded_losses(model, mi, gt, poses, begin, end, n_total):
    """What a rank computes in point-sharded mode: sums over its chunk divided by the GLOBAL N."""
    sl = lambda d: {k: v[:, begin:end] for k, v in d.items()}
    ld = O.mapping_loss(model, sl(mi), sl(gt), poses, "L1", 1.0, 0.0, 0.1, 0.15)
    scale = mdist.sharded_loss_scale(end - begin, n_total)
    return sum(ld.values()) * scale


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        # ---- point-sharded fit ---------------------------------------------------------------
        model = _model()
        mi, gt, (R, t) = synth.rgbd_batch(2001, num_kf=3, bound=BOUND, seed=4, wall_margin=0.3)
        poses = {k: (R[k], t[k]) for k in range(3)}
        b, e = mdist.shard_points(2001)
        loss = _sharded_losses(model, mi, gt, poses, b, e, 2001)
        loss.backward()
        grads = [p.grad for p in model.features]
        lt = loss.detach().clone().reshape(1)
        mdist.allreduce_sum_(grads + [lt])
        # ---- pair-sharded alignment ------------------------------------------------------------
        Rt, tt = synth.submap_layout(3, spacing=(2.0, 1.5))
        Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=3.0, trans_m=0.2)
        shapes = O.level_shapes(BOUND, 0.5, 2, 2, 4)
        subs = [O.OracleGridNet(BOUND, synth.fill_submap_from_field(shapes, BOUND, Rt[i], tt[i]), None) for i in range(3)]
        atlas = O.OracleAtlas(subs, Rp, tp)
        atlas.precompute([0])
        pairs = [(s, d) for s in range(3) for d in range(s + 1, 3)]
        mine = [p for i, p in enumerate(pairs) if mdist.pair_filter()(i, p)]
        total = torch.zeros(())
        per_pair = torch.zeros(len(pairs))
        for s, d in mine:
            Rs, ts = atlas.updated_submap_pose(s)
            Rd, td = atlas.updated_submap_pose(d)
            val = O.pairwise_loss_latent(subs[s], subs[d], atlas.coords[(s, 0)], Rs, ts, Rd, td, 0)
            per_pair[pairs.index((s, d))] = float(val.detach())
            total = total + val
        if total.requires_grad:
            total.backward()
        pg = [p.grad if p.grad is not None else torch.zeros_like(p) for p in atlas.rot + atlas.tra]
        tl = total.detach().clone().reshape(1)
        mdist.allreduce_sum_(pg + [tl, per_pair])
        # ---- submap-per-rank -------------------------------------------------------------------
        owned = mdist.submaps_for_rank(5)
        local = {i: {"id": torch.tensor(i), "rank": rank} for i in owned}
        gathered = mdist.gather_submaps_to_rank0(local, 5)
        if rank == 0:
            q.put({"grads": [g.clone() for g in grads], "loss": lt, "pose_grads": [g.clone() for g in pg], "align": tl,
                   "per_pair": per_pair.tolist(), "gathered": [int(g["id"]) for g in gathered], "owners": [g["rank"] for g in gathered], "mine": mine})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world2_gloo_partitions_match_single_process():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=500)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process references
    model = _model()
    mi, gt, (R, t) = synth.rgbd_batch(2001, num_kf=3, bound=BOUND, seed=4, wall_margin=0.3)
    poses = {k: (R[k], t[k]) for k in range(3)}
    full = sum(O.mapping_loss(model, mi, gt, poses, "L1", 1.0, 0.0, 0.1, 0.15).values())
    full.backward()
    # worker processes run torch with 2 threads, this process with all cores: reductions differ in the last bits (a
    # 4e-5 relative difference between the two was observed once), so the cross-process comparisons use 2e-4 / 1e-3;
    # the partition logic itself is checked exactly below (sum of the per-rank terms == all-reduced total)
    assert abs(float(res["loss"]) - float(full)) < 2e-4 * max(1.0, abs(float(full)))
    for a, b in zip(res["grads"], [p.grad for p in model.features]):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-7)
    Rt, tt = synth.submap_layout(3, spacing=(2.0, 1.5))
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=3.0, trans_m=0.2)
    shapes = O.level_shapes(BOUND, 0.5, 2, 2, 4)
    subs = [O.OracleGridNet(BOUND, synth.fill_submap_from_field(shapes, BOUND, Rt[i], tt[i]), None) for i in range(3)]
    atlas = O.OracleAtlas(subs, Rp, tp)
    atlas.precompute([0])
    total = 0
    per_pair = []
    for s in range(3):
        for d in range(s + 1, 3):
            Rs, ts = atlas.updated_submap_pose(s)
            Rd, td = atlas.updated_submap_pose(d)
            val = O.pairwise_loss_latent(subs[s], subs[d], atlas.coords[(s, 0)], Rs, ts, Rd, td, 0)
            per_pair.append(float(val.detach()))
            total = total + val
    total.backward()
    # the host logic under test: the all-reduced total is the sum of the per-pair terms the ranks computed
    assert abs(float(res["align"]) - sum(res["per_pair"])) < 1e-5 * abs(float(total)), (res["align"], res["per_pair"])
    assert np.allclose(res["per_pair"], per_pair, rtol=2e-4), (res["per_pair"], per_pair)
    for a, p in zip(res["pose_grads"], atlas.rot + atlas.tra):
        ref = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(a, ref, rtol=1e-3, atol=1e-5)
    assert res["gathered"] == [0, 1, 2, 3, 4] and res["owners"] == [0, 1, 0, 1, 0]
    assert res["mine"] == [(0, 1), (1, 2)]


# ------------------------------------------------------------------------------------------------
# domain-decomposed fit (miso_b200.sharded_fit): slab ownership + one-plane halos == the full gradient
# ------------------------------------------------------------------------------------------------
def _slab_worker(rank, world, port, q):
    from miso_b200 import sharded_fit as sf
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    try:
        model = _model()
        N = 3001
        mi, gt, (R, t) = synth.rgbd_batch(N, num_kf=3, bound=BOUND, seed=6, wall_margin=0.3)
        poses = {k: (R[k], t[k]) for k in range(3)}
        ids = mi["sample_frame_ids"][0, :, 0]
        zw = torch.einsum("nj,nj->n", R[ids][:, 2, :], mi["coords_frame"][0]) + t[ids][:, 2, 0]
        fine = model.features[1]
        Z = fine.shape[2]
        plane = sf.plane_of_points(zw, BOUND[2][0], BOUND[2][1], Z)
        bounds = sf.slab_bounds_from_histogram(torch.bincount(plane, minlength=Z), world)
        zb, ze = bounds[rank], bounds[rank + 1]
        own = torch.nonzero((plane >= zb) & (plane < ze))[:, 0]
        sel = lambda d: {k: v[:, own] for k, v in d.items()}
        ld = O.mapping_loss(model, sel(mi), sel(gt), poses, "L1", 1.0, 0.0, 0.1, 0.15)
        (sum(ld.values()) * mdist.sharded_loss_scale(own.numel(), N)).backward()
        gf = fine.grad.permute(0, 2, 3, 4, 1).reshape(Z, -1).contiguous()        # (Z, plane) like SlabShardedFit._flat
        touched = torch.nonzero(gf.abs().sum(1))[:, 0]
        assert touched.numel() > 0 and int(touched.min()) >= zb and int(touched.max()) <= min(ze, Z - 1)   # planes [zb, ze]
        halo = torch.zeros_like(gf[0])
        sf.exchange_halo_planes(gf[ze].clone() if ze < Z else None, halo if rank > 0 else None, rank, world)
        if rank > 0:
            gf[zb] += halo
        coarse = model.features[0].grad.clone()
        mdist.allreduce_sum_([coarse])
        # fake optimiser step on the owned planes, then the parameter halo
        pf = fine.detach().permute(0, 2, 3, 4, 1).reshape(Z, -1).clone()
        pf[zb:ze] -= 0.5 * gf[zb:ze]
        recv = torch.zeros_like(pf[0])
        sf.exchange_halo_planes_down(pf[zb].clone() if rank > 0 else None, recv if ze < Z else None, rank, world)
        q.put({"rank": rank, "bounds": bounds, "own": own.numel(), "grad_slab": gf[zb:ze].clone(), "coarse": coarse,
               "param_first_plane": pf[zb].clone(), "param_halo": recv if ze < Z else None})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_world2_gloo_slab_sharded_fit_matches_single_process():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=500) for _ in range(2)], key=lambda d: d["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    model = _model()
    N = 3001
    mi, gt, (R, t) = synth.rgbd_batch(N, num_kf=3, bound=BOUND, seed=6, wall_margin=0.3)
    sum(O.mapping_loss(model, mi, gt, {k: (R[k], t[k]) for k in range(3)}, "L1", 1.0, 0.0, 0.1, 0.15).values()).backward()
    Z = model.features[1].shape[2]
    full = model.features[1].grad.permute(0, 2, 3, 4, 1).reshape(Z, -1)
    b = res[0]["bounds"]
    assert b == res[1]["bounds"] and b[0] == 0 and b[-1] == Z and res[0]["own"] + res[1]["own"] == N
    assert min(res[0]["own"], res[1]["own"]) > 0.25 * N                     # boundaries balance the sample count
    for r in range(2):
        assert torch.allclose(res[r]["grad_slab"], full[b[r]:b[r + 1]], rtol=1e-3, atol=1e-7), r
        assert torch.allclose(res[r]["coarse"], model.features[0].grad, rtol=1e-3, atol=1e-7)
    # rank 0 received rank 1's freshly updated first plane
    assert torch.equal(res[0]["param_halo"], res[1]["param_first_plane"])
