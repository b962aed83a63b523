#!/usr/bin/env python
"""BASELINE.json configs[3]: Newer-College-quad-shaped single grid (bound [[-45,45],[-45,45],[-5,15]], levels
(1,4,20,90,90) + (1,4,100,450,450) = 324 MB fine level, decoder_quad-shaped MLP), 2^22 LiDAR-sampled points per
step, POINT-SHARDED over the ranks with an NCCL all_reduce of the dense grid gradients before the (replicated)
Adam step.  Loss per ncd_quad.yaml:42-46: L2 sdf + 0.5 free-space, trunc 0.5.

    python benchmarks/ncd_point_sharded.py                       # 1 GPU
    torchrun --nproc-per-node N benchmarks/ncd_point_sharded.py   # N GPUs, also checks N-GPU == 1-GPU parameters
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from miso_b200 import dist as mdist, synth  # noqa: E402
from miso_b200.loss import MisoLossMapping  # noqa: E402
from miso_b200.models import GridNet  # noqa: E402
from miso_b200.trainer import GridTrainer  # noqa: E402

N_TOTAL = 1 << 22
NUM_KF = 8


def build(device, seed=0):
    cfg = synth.model_cfg(synth.NCD_QUAD_BOUND, base_cell_size=1.0, per_level_scale=5, num_poses=NUM_KF)
    net = GridNet(cfg, device=device)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for lvl in net.features:
            lvl.feature.copy_((torch.randn(lvl.feature.shape, generator=g) * 1e-2).to(device))
    net.decoder.load_state_dict(synth.decoder_weights(8, seed=0))
    return net


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    mi, gt, (R, t) = synth.lidar_batch(N_TOTAL, num_kf=NUM_KF, seed=3)
    b, e = mdist.shard_points(N_TOTAL, rank, world)
    sl = lambda d: {k: v[:, b:e].contiguous().to(device) for k, v in d.items()}
    L = MisoLossMapping(loss_type="L2", weight_sdf=1.0, weight_eik=0.0, weight_fs=0.5, trunc_dist=0.5)

    def run(n_steps, sharded):
        net = build(device)
        for k in range(NUM_KF):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        net.lock_pose()
        tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, net, L, None, device=device)
        if sharded:
            dmi, dgt = sl(mi), sl(gt)
            step = lambda: tr.train_step(dmi, dgt, n_total=N_TOTAL, allreduce=mdist.allreduce_sum_ if world > 1 else None)
        else:
            dmi = {k: v.to(device) for k, v in mi.items()}
            dgt = {k: v.to(device) for k, v in gt.items()}
            step = lambda: tr.train_step(dmi, dgt)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        from miso_b200 import loss as mloss
        mloss.PROFILE_EVENTS = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            terms = step()
        e1.record()
        torch.cuda.synchronize()
        run.kernel_ms = sum(a.elapsed_time(b_) for a, b_ in mloss.PROFILE_EVENTS) / max(len(mloss.PROFILE_EVENTS), 1)
        mloss.PROFILE_EVENTS = None
        ms = e0.elapsed_time(e1) / n_steps
        if world > 1:
            tms = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms)
        return net, ms, terms

    net_s, ms_s, terms_s = run(10, sharded=True)
    out = {"workload": "NCD quad grid, 2^22 LiDAR points/step, point-sharded + all_reduce of grid gradients",
           "n_gpus": world, "ms_per_step": ms_s, "points_per_s": N_TOTAL / (ms_s * 1e-3),
           "loss_terms": [float(v) for v in terms_s.tolist()],
           "allreduce_bytes_per_step": sum(p.numel() * 4 for p in net_s.level_tensors()),
           "mapping_kernel_ms_per_rank": run.kernel_ms,
           # no eikonal term in this config (ncd_quad.yaml:42-46): same 545 B/point model as bench.py
           "kernel_roofline_frac": 545 * (e - b) / (run.kernel_ms * 1e-3) / 1e9 / 6535.7}
    if world > 1:
        # parity of the sharded run against the same 13 steps on one GPU (rank 0 recomputes unsharded)
        params_s = [p.detach().clone() for p in net_s.level_tensors()]
        del net_s
        torch.cuda.empty_cache()
        net_1, ms_1, terms_1 = run(10, sharded=False)
        rel = [float((a - b_).norm() / b_.norm()) for a, b_ in zip(params_s, net_1.level_tensors())]
        out["single_gpu_ms_per_step"] = ms_1
        out["speedup_vs_1gpu"] = ms_1 / ms_s
        out["param_rel_err_vs_1gpu"] = rel
        out["loss_rel_err_vs_1gpu"] = abs(float(terms_s[3]) - float(terms_1[3])) / abs(float(terms_1[3]))
        assert max(rel) < 1e-4, rel
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
