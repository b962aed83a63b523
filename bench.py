#!/usr/bin/env python
"""bench.py -- SDF train pts/s of the fused MISO hot path on N B200s (driver contract in the task).

Workload (BASELINE.json configs[1], `build_submaps`): one ScanNet-submap-shaped GridNet per GPU
(bound [[-10,10],[-5,5],[-10,10]], 2 levels: coarse (1,4,40,20,40), fine (1,4,200,100,200), decoder
8->64->64->1 fixed), 2^20 synthetic RGB-D-sampled points per iteration.  A "step" is one mapping
iteration of the reference's hot loop (grid_opt/trainer.py:209-217): loss.compute (L1 sdf + 0.1
free-space + 0.5 second-order eikonal) -> backward -> Adam.step over both grid levels.
N GPUs = N independent submaps (weak scaling, no data-path collective).

  value     device-resident throughput (inputs already in HBM), CUDA events, max over ranks
  e2e       same step through miso_b200.trainer with HOST (pinned) buffers: H2D of the batch and a D2H
            read of the loss every step, inside the timed region
  roofline  fused mapping kernel: algorithmic bytes/launch over its measured duration vs the measured
            HBM peak in MEASURED_PEAKS.json
  cpu_baseline / --impl reference: the oracle port of the reference's torch CPU path (the reference is
            Python and /root/reference is absent on the GPU box) on a bounded sample, all host cores.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

N_POINTS = 1 << 20
NUM_KF = 49
NUM_HOST_BATCHES = 4
LOSS_CFG = dict(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                grad_method="autograd", eik_trunc_dist=None)
# algorithmic bytes per point of the fused step (DESIGN.md "bytes per unit"): coords 12 + frame id 8 + sdf 4 +
# valid 1 + sign 4 + weight 4 + corner gather 2 levels x 8 x 16 B + gradient scatter 2 x 8 x 16 B
BYTES_PER_POINT = 12 + 8 + 4 + 1 + 4 + 4 + 256 + 256
SURVEY_BYTES_PER_POINT = 1068   # SURVEY.md section 8d two-pass figure (fwd + bwd re-gather + eikonal scatter)
FLOPS_PER_POINT = 2 * (2 * (8 * 64 + 64 * 64) + 64)  # MLP forward + Jacobian backward, FMA = 2 flops
if os.environ.get("MISO_MLP", "").lower().startswith("s"):
    KERNEL_NAME = "mapping_step_kernel<2,4> (SIMT decoder) (+finalize)"
elif os.environ.get("MISO_TC", "") == "1":
    KERNEL_NAME = "mapping_step_tc_kernel<2,4> (one thread per point) (+finalize)"
else:
    KERNEL_NAME = "mapping_step_tc2_kernel<2,4,%s,%s> (two threads per point) (+finalize)" % (
        os.environ.get("MISO_TC2_GROUPS", "4"), "unpaired" if os.environ.get("MISO_PAIR", "1") == "0" else "paired")


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, taken from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json is written by tools/ncu_summary.py from
    the .ncu-rep; a profiler cannot run inside the timed region).  None when no capture of that workload exists."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as fh:
        d = json.load(fh).get(key)
    return (d["dram_bytes_per_launch"], d["source"]) if d else (None, None)


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(world):
    """`config` of the JSON line: the same dict for the product arm and for --impl reference."""
    return {"workload": "build_submaps: 1 ScanNet-submap GridNet per GPU, 2 levels (40x20x40, 200x100x200) x C4, "
                        "decoder 8-64-64-1 fixed, 2^20 RGB-D-sampled pts/iter, L1 sdf + 0.1 free-space + 0.5 "
                        "second-order eikonal, Adam joint",
            "points_per_step_per_gpu": N_POINTS, "submaps": world, "parallelism": f"submap-per-gpu x{world}",
            "l2_policy": f"{NUM_HOST_BATCHES} distinct batches cycled; grids+grads+Adam state+batch "
                         "(~360 MB/step touched) exceed the 126 MB L2"}


def host_batch(seed0, b, poses=None):
    """Batch `b` of rank `seed0` (CPU tensors, reference dict layout) -- the GPU arm and the CPU arm draw from here, so
    rank 0's batch 0 is the same tensor set on both."""
    from miso_b200 import synth
    poses = poses if poses is not None else synth.keyframe_poses(NUM_KF, synth.SCANNET_SUBMAP_BOUND, seed=55 + seed0)
    mi, gt, _ = synth.rgbd_batch(N_POINTS, num_kf=NUM_KF, seed=1000 * seed0 + b, poses=poses)
    return mi, gt, poses


def make_host_batches(seed0):
    batches, poses = [], None
    for b in range(NUM_HOST_BATCHES):
        mi, gt, poses = host_batch(seed0, b, poses)
        mi = {k: v.pin_memory() for k, v in mi.items()}
        gt = {k: v.pin_memory() for k, v in gt.items()}
        batches.append((mi, gt))
    return batches, poses


def initial_features(shapes, seed):
    """Seeded N(0, 1e-2) level tensors: the same values for the product model and the CPU arm's model."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(tuple(s), generator=g) * 1e-2 for s in shapes]


def build_model(device, poses, seed):
    from miso_b200 import synth
    from miso_b200.models import GridNet
    cfg = synth.model_cfg(synth.SCANNET_SUBMAP_BOUND, num_poses=NUM_KF)
    net = GridNet(cfg, device=device)
    with torch.no_grad():
        for lvl, f in zip(net.features, initial_features([l.feature.shape for l in net.features], seed)):
            lvl.feature.copy_(f.to(device))
    net.decoder.load_state_dict(synth.decoder_weights(8, seed=0))
    R, t = poses
    for k in range(R.shape[0]):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    return net


def run_ours(args):
    import torch.distributed as dist
    from miso_b200 import _lib, loss as mloss
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import GridTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_node = None
    if world > 1 and os.environ.get("MISO_NUMA_BIND", "1") != "0":
        from miso_b200 import dist as mdist
        numa_node = mdist.bind_to_gpu_numa_node(local)   # before any pinned allocation (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _lib.load()

    batches, poses = make_host_batches(rank)
    net = build_model(device, poses, seed=rank)
    L = MisoLossMapping(**LOSS_CFG)
    trainer = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint", "cuda_graph": True}, net, L,
                          lambda e: batches[e % NUM_HOST_BATCHES], device=device)
    dev_batches = [({k: v.to(device) for k, v in mi.items()}, {k: v.to(device) for k, v in gt.items()})
                   for mi, gt in batches]
    if os.environ.get("MISO_PRESORT", "0") == "1":   # experiment: Morton-ordered batches (not the default)
        from miso_b200.sorting import morton_order
        R, t = net.all_kf_poses()
        sorted_batches = []
        for mi, gt in dev_batches:
            perm = morton_order(mi["coords_frame"][0], mi["sample_frame_ids"][0, :, 0], R, t, net._bound_host)
            sorted_batches.append(({k: v[:, perm].contiguous() for k, v in mi.items()},
                                   {k: v[:, perm].contiguous() for k, v in gt.items()}))
        dev_batches = sorted_batches

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident value ----------------
    # public API of the device-resident loop: GridTrainer.graphed_train_step -- one CUDA-graph launch per step (count
    # + fused step + finalize + one Adam sweep per level with device-side step counters).  The first two calls per
    # batch buffer run eagerly / capture, so the warm-up covers at least 2 rounds over the 4 resident batches.
    first_terms = None
    for i in range(max(args.warmup, 2 * NUM_HOST_BATCHES)):
        terms = trainer.graphed_train_step(*dev_batches[i % NUM_HOST_BATCHES])
        if i == 0:
            first_terms = [float(v) for v in terms.tolist()]   # loss of batch 0 at the initial parameters
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(args.steps):
        last = trainer.graphed_train_step(*dev_batches[i % NUM_HOST_BATCHES])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    final_loss = [float(v) for v in last.tolist()]
    # the same steps issued launch by launch (eager), with CUDA events around every fused-step launch on its stream:
    # the per-kernel duration behind `roofline` (events cannot bracket a node inside a graph replay) and the launch count
    mloss.PROFILE_EVENTS = []
    launches0 = _lib.LAUNCHES["total"]
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for i in range(args.steps):
        trainer.train_step(*dev_batches[i % NUM_HOST_BATCHES])
    g1.record()
    barrier()
    launches = _lib.LAUNCHES["total"] - launches0
    ms_total_eager = max_over_ranks(g0.elapsed_time(g1))
    kern_ms = [a.elapsed_time(b) for a, b in mloss.PROFILE_EVENTS]
    mloss.PROFILE_EVENTS = None

    # ---------------- end-to-end through the trainer API with host buffers ----------------
    wu = max(args.warmup, 4)      # each of the two staging slots is used eagerly once, then captured, before timing
    loss_host = torch.zeros(args.steps + wu, 4).pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for d in batches[0] for v in d.values())

    def e2e_loop(n, offset):
        # public API: GridTrainer.train_host_batches -- pinned host batches in, H2D on a copy stream
        # overlapping the previous step, loss terms read back D2H every step
        trainer.train_host_batches((batches[(offset + i) % NUM_HOST_BATCHES] for i in range(n)),
                                   loss_sink=loss_host[offset:offset + n])

    e2e_loop(wu, 0)
    barrier()
    gc.collect()
    gc.disable()   # the e2e loops are paced by the host thread: keep collector pauses out of the timed region
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps, wu)
    t1.record()
    barrier()
    gc.enable()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    assert torch.isfinite(loss_host[wu:]).all(), "non-finite loss in the e2e run"
    e2e_loss_ref_format = loss_host[wu + args.steps - 1].clone()

    # same loop with the compact wire format (int16 ids, masks rebuilt on the device): 18 B/point over PCIe
    from miso_b200.trainer import CompactBatch
    compact = [CompactBatch.from_reference(mi, gt, LOSS_CFG["trunc_dist"]) for mi, gt in batches]
    h2d_compact = sum(v.numel() * v.element_size() for v in compact[0].tensors().values())

    def e2e_compact_loop(n, offset):
        trainer.train_host_batches((compact[(offset + i) % NUM_HOST_BATCHES] for i in range(n)),
                                   loss_sink=loss_host[offset:offset + n])

    e2e_compact_loop(wu, 0)
    barrier()
    gc.collect()
    gc.disable()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    e2e_compact_loop(args.steps, wu)
    c1.record()
    barrier()
    gc.enable()
    e2e_compact_ms = max_over_ranks(c0.elapsed_time(c1))
    clocks = sampler.stop() if sampler else None   # sampled across all timed regions (value + e2e + e2e_compact)
    assert torch.isfinite(loss_host[wu:]).all(), "non-finite loss in the compact e2e run"

    align, ncd, torch_gpu, fd, wdec = None, None, None, None, None
    if not args.no_extras:
        del dev_batches
        torch.cuda.empty_cache()
        ncd = bench_ncd(device, steps=max(100, min(args.steps, 200)) if world == 1 else 30, world=world, rank=rank)
        if world == 1:
            torch_gpu = torch_gpu_arm(device)
            fd = bench_fd(device)
            wdec = bench_trainable_decoder(device)
        align = bench_align(device, iters=10, warmup=2, world=world, rank=rank)
        if world > 1:
            # pair-sharded: an iteration ends when the slowest rank is done
            for lv in align.values():
                lv["ms_per_iter"] = max_over_ranks(lv["ms_per_iter"])
                lv["iters_per_s"] = 1e3 / lv["ms_per_iter"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_hbm_peak()
    ms_step = ms_total / args.steps
    value = world * N_POINTS / (ms_step * 1e-3)
    kms = float(np.mean(kern_ms))
    achieved = BYTES_PER_POINT * N_POINTS / (kms * 1e-3) / 1e9
    line = {
        "metric": "SDF train pts/s (grid+MLP fwd/bwd/eikonal)", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_step_eager": ms_total_eager / args.steps,
        "step_issue": "one CUDA-graph launch per step (GridTrainer.graphed_train_step); ms_per_step_eager = the same steps "
                      "issued launch by launch from Python", "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "clocks": clocks,
        "e2e": {"value": world * N_POINTS / (e2e_ms / args.steps * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 16, "ms_per_step": e2e_ms / args.steps},
        "e2e_compact": {"value": world * N_POINTS / (e2e_compact_ms / args.steps * 1e-3), "unit": "points/s",
                        "h2d_bytes_per_step": h2d_compact, "d2h_bytes_per_step": 16,
                        "ms_per_step": e2e_compact_ms / args.steps,
                        "format": "CompactBatch: coords f32x3 + keyframe id int16 + sdf f32 (+weights when not all "
                                  "ones); sdf_valid / sdf_signs / int64 ids rebuilt on the device by miso_expand_batch "
                                  "(the reference's datasets define them as functions of sdf and trunc_dist)"},
        "host_numa_node_rank0": numa_node,
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": KERNEL_NAME, "achieved": achieved,
                     "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("scannet_2p20")[0],
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this "
                                       "kernel at this size (%s); below the algorithmic bytes because the 64 MB grid and "
                                       "its gradient stay L2-resident" % ncu_traffic("scannet_2p20")[1],
                     "bytes_per_point": BYTES_PER_POINT, "survey_bytes_per_point": SURVEY_BYTES_PER_POINT,
                     "kernel_ms": kms, "kernel_share_of_step": kms / ms_step,
                     "kernel_timing": "CUDA events around every fused-step launch of the eager pass that follows the timed "
                                      "(graph-replayed) region: same batches, same kernels, same stream",
                     "fp32_equiv_tflops_achieved": FLOPS_PER_POINT * N_POINTS / (kms * 1e-3) / 1e12,
                     "decoder": os.environ.get("MISO_MLP", "tcgen05 3xTF32"),
                     "floors_ms": {"red_v4_scatter_only": 0.121, "gather_only": 0.041,
                                   "source": "profiles/r01_scatter_probe.json (benchmarks/scatter_probe.py, same batch)"}},
        "final_loss_terms": final_loss,
        "extra": {"ncd": ncd, "torch_gpu_baseline": torch_gpu, "finite_difference_step": fd, "trainable_decoder_step": wdec, "align": align, "align_workload": "16 ScanNet-shaped submaps (4x4 floor plan, 40 % overlap), 120 pairs "
                  "(sharded round-robin over ranks, pose-gradient all_reduce), latent L2 loss, Adam lr 1e-2; level 0: "
                  "<= 32 k samples/pair, level 1: <= 4 M samples/pair"},
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_reference_arm(steps=2, warmup=1)
        # same batch, same initial parameters on both arms: the CPU arm's first step doubles as a full-size
        # (2^20-point) value check of the kernel's loss terms (the CPU arm is the checker here, never the product)
        want = cb.pop("first_step_terms")
        got = {"sdf_L1": first_terms[0], "free_space": first_terms[1], "eik": first_terms[2], "total": first_terms[3]}
        rel = {k: abs(got[k] - want[k]) / max(abs(want[k]), 1e-12) for k in want}
        line["parity_check"] = {"what": "loss terms of rank 0's batch 0 (2^20 points) at the initial parameters: fused CUDA "
                                        "step vs the CPU arm's first step on the same tensors", "cuda": got, "cpu": want,
                                "rel_err": rel, "tolerance": 1e-5, "ok": max(rel.values()) <= 1e-5}
        assert line["parity_check"]["ok"], line["parity_check"]
        line["cpu_baseline"] = cb
        if not args.no_extras:
            line["extra"]["align_cpu_baseline"] = cpu_align_arm(iters=1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_reference_arm(steps, warmup, sample_points=N_POINTS):
    """The reference's CPU path for the SAME step on the SAME inputs (oracle port: identical torch ops to
    grid_opt/loss.py:754-813 + torch.optim.Adam; the second-order eikonal goes through the gather restatement
    because F.grid_sample has no double backward on CPU): rank 0's batch 0, all 2^20
    points, the full-size grids, the product arm's initial parameters.  Bounded by the step count, not by
    shrinking the workload."""
    from miso_b200 import synth
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count())
    shapes = O.level_shapes(synth.SCANNET_SUBMAP_BOUND, 0.5, 5, 2, 4)
    feats = initial_features(shapes, 0)
    dec = O.make_decoder(8)
    dec.load_state_dict({k.replace("network.", ""): v for k, v in synth.decoder_weights(8).items()})
    model = O.OracleGridNet(synth.SCANNET_SUBMAP_BOUND, feats, dec, second_order=True)
    mi, gt, (R, t) = host_batch(0, 0)
    if sample_points < N_POINTS:
        mi = {k: v[:, :sample_points] for k, v in mi.items()}
        gt = {k: v[:, :sample_points] for k, v in gt.items()}
    poses = {k: (R[k], t[k]) for k in range(R.shape[0])}
    opt = torch.optim.Adam(list(model.features.parameters()), lr=1e-3)
    times, first = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        ld = O.mapping_loss(model, mi, gt, poses, LOSS_CFG["loss_type"], LOSS_CFG["weight_sdf"], LOSS_CFG["weight_eik"],
                            LOSS_CFG["weight_fs"], LOSS_CFG["trunc_dist"], grad_method="autograd", eik_trunc_dist=None)
        total = sum(ld.values())
        total.backward()
        opt.step()
        if i == 0:   # unweighted terms + weighted total, the layout of the kernel's loss_out
            first = {"sdf_L1": float(ld["sdf_L1"]) / LOSS_CFG["weight_sdf"], "free_space": float(ld["free_space"]) / LOSS_CFG["weight_fs"],
                     "eik": float(ld["eik"]) / LOSS_CFG["weight_eik"], "total": float(total)}
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    return {"value": sample_points / sec, "unit": "points/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"all {sample_points} points/step of rank 0's batch 0 on the full-size grids (same config as the "
                      f"product arm), {steps} timed steps (+{warmup} warm-up), loss+backward+Adam; Adam sweeps the full "
                      "dense grids as in the reference",
            "sec_per_step": sec, "first_step_terms": first}


# ------------------------------------------------------------------------------------------------
# second half of the metric: latent-space alignment iterations / s (BASELINE.json configs[2])
# ------------------------------------------------------------------------------------------------
ALIGN_SUBMAPS = 16


def build_align_atlas(device, small=False):
    """16 ScanNet-submap-shaped GridNets tiled 4x4 with >= 30 % overlap, grids sampled from one smooth latent
    field at the true poses, then perturbed by ~10 deg / 0.5 m (demo/align_submaps.py:267-273)."""
    from miso_b200 import synth
    from miso_b200.models import GridAtlas
    bound = synth.SCANNET_SUBMAP_BOUND
    cfg = synth.model_cfg(bound, num_poses=1) if not small else synth.model_cfg(bound, base_cell_size=1.0, per_level_scale=2, num_poses=1)
    Rt, tt = synth.submap_layout(ALIGN_SUBMAPS, spacing=(12.0, 12.0))
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=10.0, trans_m=0.5)
    atlas = GridAtlas(cfg, device=device)
    for i in range(ALIGN_SUBMAPS):
        atlas.add_submap(torch.tensor(bound), Rp[i], tp[i])
        sm = atlas.get_submap(i)
        shapes = [tuple(f.feature.shape) for f in sm.features]
        feats = synth.fill_submap_from_field(shapes, bound, Rt[i], tt[i], device=device)
        with torch.no_grad():
            for l in range(2):
                sm.features[l].feature.copy_(feats[l])
        sm.lock_feature()
    return atlas


def bench_align(device, iters=10, warmup=2, world=1, rank=0):
    """align iters/s: one iteration = intersection test + latent loss of every overlapping pair + backward +
    Adam step on the 15 free submap poses (generic_align_multiple_submaps body, align/base.py:127-159)."""
    import torch.optim as optim
    from miso_b200 import dist as mdist
    from miso_b200.align import AlignBatch, FusedPoseAligner
    atlas = build_align_atlas(device)
    atlas.precompute_coordinates_for_alignment()
    pairs = [(s, d) for s in range(ALIGN_SUBMAPS) for d in range(s + 1, ALIGN_SUBMAPS)]
    out = {}
    for level in (0, 1):
        if world > 1:
            # cost-balanced ownership: samples of the pair if the submaps intersect at the initial poses, else 0
            # (same deterministic assignment on every rank)
            probe = AlignBatch(atlas, pairs, level, check_intersection=True, cache_src_features=False)
            probe.update_intersections(probe.pair_poses())
            en = probe.enabled[:len(pairs)].tolist()
            # + the pair's share of the per-iteration intersection test (one transform + bound test per finest-level
            # vertex of the source, ~5 % of the cost of an alignment sample)
            costs = [probe._coords[s].shape[0] * (1 if e else 0) + 0.05 * probe._verts[s].shape[0]
                     for (s, d), e in zip(pairs, en)]
            owner = mdist.balanced_pair_owner(costs, world)
            mine = [p for p, o in zip(pairs, owner) if o == rank]
            del probe
        else:
            mine = pairs
        batch = AlignBatch(atlas, mine, level, check_intersection=True)
        # whole iteration = compose poses, intersection test, alignment kernel, pose gradients, Adam: five launches
        # (csrc/poseopt.cu), captured as one CUDA graph; with world > 1 the NCCL all_reduce of the (S,6) pose gradients
        # is a node of the same graph (MISO_ALIGN_GRAPH_NCCL=0 keeps the multi-GPU iteration eager)
        use_graph = world == 1 or os.environ.get("MISO_ALIGN_GRAPH_NCCL", "1") != "0"
        aligner = FusedPoseAligner(batch, lr=1e-2, align_weight=3000.0, max_iters=4 * (iters + warmup) + 64,
                                   allreduce=mdist.allreduce_sum_ if world > 1 else None)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 3)):
                aligner.iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = None
        if use_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                aligner.iteration()
            graph.replay()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            if graph is not None:
                graph.replay()
            else:
                aligner.iteration()
        e1.record()
        torch.cuda.synchronize()
        last = aligner.loss_hist[int(aligner.iter_counter.item()) - 1].clone()
        if world > 1:
            mdist.allreduce_sum_([last])
        ms = e0.elapsed_time(e1) / iters
        n_on = int(batch.enabled[:len(mine)].sum().item()) if mine else 0
        samples = sum(batch._coords[s].shape[0] for (s, d), en in zip(mine, batch.enabled.tolist()) if en)
        bytes_pp = 12 + (level + 1) * 16 + (level + 1) * 8 * 16      # coords + cached src feats + dst corners
        out[f"level{level}"] = {"iters_per_s": 1e3 / ms, "ms_per_iter": ms, "pairs": len(mine), "pairs_overlapping": n_on,
                                "samples_per_iter": samples, "loss": float(last), "cuda_graph": graph is not None, "launches_per_iter": 5,
                                "hbm_gbs_algorithmic": samples * bytes_pp / (ms * 1e-3) / 1e9}
    return out


def cpu_align_arm(iters=1):
    """Oracle port of the reference's alignment iteration on the host (level 0, same 16-submap layout)."""
    from miso_b200 import synth
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count())
    bound = synth.SCANNET_SUBMAP_BOUND
    shapes = O.level_shapes(bound, 0.5, 5, 2, 4)
    Rt, tt = synth.submap_layout(ALIGN_SUBMAPS, spacing=(12.0, 12.0))
    Rp, tp = synth.perturb_poses(Rt, tt, rot_deg=10.0, trans_m=0.5)
    subs = []
    for i in range(ALIGN_SUBMAPS):
        f0 = synth.fill_submap_from_field(shapes[:1], bound, Rt[i], tt[i])[0]
        subs.append(O.OracleGridNet(bound, [f0, torch.zeros(1, 4, 2, 2, 2)], None))   # level 0 only needs the coarse grid
    atlas = O.OracleAtlas(subs, Rp, tp)
    for i, sm in enumerate(subs):
        coords = O.vertex_positions(sm.features[0].shape, bound)
        atlas.coords[(i, 0)] = coords
    # the intersection test uses the finest level's vertices in the reference; bound it with the coarse ones here
    t0 = time.perf_counter()
    O.align_multiple_submaps(atlas, level=0, num_iters=iters - 1, lr=1e-2, check_intersection=False)
    sec = (time.perf_counter() - t0) / iters
    return {"iters_per_s": 1.0 / sec, "sec_per_iter": sec, "cores": os.cpu_count(), "kind": "port",
            "sample": "level 0 only (M0 = 32000 samples/pair, 120 pairs, intersection test skipped), 1 iteration"}


# ------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: Newer-College-quad grid, 2^22 LiDAR points per step (the size the 0.6 target is quoted on)
# ------------------------------------------------------------------------------------------------
NCD_POINTS = 1 << 22
NCD_KF = 8
NCD_LOSS = dict(loss_type="L2", weight_sdf=1.0, weight_eik=0.0, weight_fs=0.5, trunc_dist=0.5)   # ncd_quad.yaml:42-46


def build_ncd_model(device, poses):
    from miso_b200 import synth
    from miso_b200.models import GridNet
    cfg = synth.model_cfg(synth.NCD_QUAD_BOUND, base_cell_size=1.0, per_level_scale=5, num_poses=NCD_KF)
    net = GridNet(cfg, device=device)
    with torch.no_grad():
        for lvl, f in zip(net.features, initial_features([l.feature.shape for l in net.features], 0)):
            lvl.feature.copy_(f.to(device))
    net.decoder.load_state_dict(synth.decoder_weights(8, seed=0))
    R, t = poses
    for k in range(NCD_KF):
        net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
    net.unlock_feature()
    net.lock_pose()
    return net


def bench_ncd(device, steps=100, warmup=5, world=1, rank=0, modes=("allreduce", "slab"), halo="auto"):
    """Fused step + Adam on the NCD quad grid (levels 20x90x90 + 100x450x450 x C4: the 324 MB fine level, its gradient
    and the Adam moments are NOT L2-resident, so dram traffic is meaningful here), 2^22 LiDAR-sampled points per step.
    One GPU: the whole batch.  N GPUs (strong scaling, same global batch): (a) the north star's split -- contiguous
    point chunks + all_reduce of the dense grid gradients + replicated Adam; (b) the domain-decomposed split of
    miso_b200.sharded_fit -- z-slabs of the fine level, one-plane halos.  Every rank also times the single-GPU step, so
    the speed-ups are measured inside one run; parameters after the timed steps are compared with the single-GPU run."""
    import torch.distributed as dist
    from miso_b200 import dist as mdist, loss as mloss, synth
    from miso_b200.loss import MisoLossMapping
    from miso_b200.sharded_fit import SlabShardedFit
    from miso_b200.trainer import GridTrainer
    mi, gt, poses = synth.lidar_batch(NCD_POINTS, num_kf=NCD_KF, seed=3)
    dmi = {k: v.to(device) for k, v in mi.items()}
    dgt = {k: v.to(device) for k, v in gt.items()}

    def sync_max(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step_fn, n):
        terms = None
        for _ in range(warmup):
            terms = step_fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if n == 0:
            return 0.0, terms
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            terms = step_fn()
        e1.record()
        torch.cuda.synchronize()
        return sync_max(e0.elapsed_time(e1) / n), terms

    # ---- one GPU, whole batch ----
    net1 = build_ncd_model(device, poses)
    tr1 = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, net1, MisoLossMapping(**NCD_LOSS), None,
                      device=device)
    mloss.PROFILE_EVENTS = []
    ms1, terms1 = timed(lambda: tr1.train_step(dmi, dgt), steps)
    kms = float(np.mean([a.elapsed_time(b) for a, b in mloss.PROFILE_EVENTS[warmup:]]))
    mloss.PROFILE_EVENTS = None
    peak, peak_src = measured_hbm_peak()
    achieved = BYTES_PER_POINT * NCD_POINTS / (kms * 1e-3) / 1e9
    traffic, tsrc = ncu_traffic("ncd_2p22")
    out = {"workload": "Newer-College-quad single grid (20x90x90 + 100x450x450 x C4, decoder 8-64-64-1 fixed), 2^22 "
                       "LiDAR-sampled pts/step, L2 sdf + 0.5 free-space (trunc 0.5), Adam joint",
           "steps": steps, "ms_per_step": ms1, "points_per_s": NCD_POINTS / (ms1 * 1e-3),
           "loss_terms": [float(v) for v in terms1.tolist()],
           "roofline": {"bound": "hbm", "kernel": KERNEL_NAME, "achieved": achieved, "peak": peak, "peak_source": peak_src,
                        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc,
                        "bytes_per_point": BYTES_PER_POINT, "kernel_ms": kms, "kernel_share_of_step": kms / ms1}}
    if world > 1:
        ref_params = [p.detach().clone() for p in net1.level_tensors()]
        del tr1, net1
        torch.cuda.empty_cache()
        total_steps = warmup + steps
        rel_a = [0.0]
        # ---- (a) point chunks + dense-gradient all_reduce + replicated Adam ----
        if "allreduce" in modes:
            b0, b1 = mdist.shard_points(NCD_POINTS, rank, world)
            smi = {k: v[:, b0:b1].contiguous() for k, v in dmi.items()}
            sgt = {k: v[:, b0:b1].contiguous() for k, v in dgt.items()}
            net_a = build_ncd_model(device, poses)
            tr_a = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, net_a, MisoLossMapping(**NCD_LOSS), None,
                               device=device)
            ms_a, terms_a = timed(lambda: tr_a.train_step(smi, sgt, n_total=NCD_POINTS, allreduce=mdist.allreduce_sum_), steps)
            rel_a = [float((a - b).norm() / b.norm()) for a, b in zip(net_a.level_tensors(), ref_params)]
            out["point_sharded_allreduce"] = {
                "ms_per_step": ms_a, "points_per_s": NCD_POINTS / (ms_a * 1e-3), "speedup_vs_1gpu": ms1 / ms_a,
                "collective": "ncclAllReduce(sum, f32) of both grid levels' dense gradients",
                "collective_bytes_per_step": sum(p.numel() * 4 for p in net_a.level_tensors()),
                "param_rel_err_vs_1gpu": rel_a, "loss_rel_err_vs_1gpu": abs(float(terms_a[3]) - float(terms1[3])) / abs(float(terms1[3]))}
            del tr_a, net_a, smi, sgt
            torch.cuda.empty_cache()
        # ---- (b) z-slabs of the fine level, one-plane halos ----
        net_b = build_ncd_model(device, poses)
        fit = SlabShardedFit(net_b, MisoLossMapping(**NCD_LOSS), lr=1e-3, halo=halo)
        bounds = fit.calibrate(dmi)
        ms_b_eager, _ = timed(lambda: fit.step(dmi, dgt), 0)        # warm-up steps only, eager
        replay = fit.graphed_step(dmi, dgt, prefetch_same=True)       # + 1 eager step (the captures themselves run nothing)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_b = steps - 1
        torch.cuda.synchronize()
        dist.barrier()
        e0.record()
        for _ in range(n_b):
            terms_b = replay()
        e1.record()
        torch.cuda.synchronize()
        ms_b = sync_max(e0.elapsed_time(e1) / n_b)
        own = int(fit._bufs["count"].item())
        slab_axis = "xyz"[fit.axis]
        fit.gather_model()
        fit.restore_layout()
        rel_b = [float((a - b).norm() / b.norm()) for a, b in zip(net_b.level_tensors(), ref_params)]
        cnt = torch.tensor([own], dtype=torch.float64, device=device)
        dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
        out["slab_sharded"] = {
            "ms_per_step": ms_b, "points_per_s": NCD_POINTS / (ms_b * 1e-3), "speedup_vs_1gpu": ms1 / ms_b,
            "cuda_graph": True, "slab_axis": slab_axis, "slab_bounds_planes": bounds, "max_samples_per_rank": int(cnt.item()),
            "load_imbalance": float(cnt.item()) * world / NCD_POINTS,
            "halo": "p2p" if fit.p2p else "nccl",
            "selection": "the next step's slab selection runs on a second stream behind the step kernel (two compaction buffer sets)",
            "collective": ("boundary-plane Adam over NVLink peer memory (miso_adam_step_halo: peer loads of the neighbour's "
                           "gradient plane, peer stores of the new parameter plane), interior Adam on a second stream, "
                           if fit.p2p else "NCCL P2P halo: one plane of fine-level gradients up + one plane of parameters down "
                           "per neighbour, ") + "ncclAllReduce of the coarse level's gradient and of the 4 loss terms",
            "collective_bytes_per_step": 2 * fit.plane_elems * 4 + net_b.level_tensors()[0].numel() * 4 + 16,
            "param_rel_err_vs_1gpu": rel_b, "loss_rel_err_vs_1gpu": abs(float(terms_b[3]) - float(terms1[3])) / abs(float(terms1[3]))}
        assert max(rel_a) < 1e-4 and max(rel_b) < 1e-4, (rel_a, rel_b)
        del fit, net_b
    torch.cuda.empty_cache()
    return out


def torch_gpu_arm(device, steps=5, warmup=2, trainable_decoder=False):
    """The reference's GPU op sequence for the same headline step on this B200, as the kernel-for-kernel bar (SURVEY.md
    section 2a / 8d): ATen grid_sampler_3d (+ its backward) per level, the reference's own double-backward extension
    (oracle/_ref/gridsample_grad2.so, built unmodified by oracle/build_ref.py) for the eikonal term, cuBLAS Linear
    layers, torch.optim.Adam -- the oracle's restatement of loss.py:754-813 moved to cuda.  When the extension was not
    built the second-order term falls back to the pure-torch gather sampler (slower; flagged in `second_order`)."""
    from miso_b200 import synth
    from oracle import oracle as O
    from oracle import ref_gpu
    shapes = O.level_shapes(synth.SCANNET_SUBMAP_BOUND, 0.5, 5, 2, 4)
    dec = O.make_decoder(8)
    dec.load_state_dict({k.replace("network.", ""): v for k, v in synth.decoder_weights(8).items()})
    mode = "plugin" if ref_gpu.available() else True
    model = O.OracleGridNet(synth.SCANNET_SUBMAP_BOUND, initial_features(shapes, 0), dec, second_order=mode).to(device)
    mi, gt, (R, t) = host_batch(0, 0)
    mi = {k: v.to(device) for k, v in mi.items()}
    gt = {k: v.to(device) for k, v in gt.items()}
    poses = {k: (R[k].to(device), t[k].to(device)) for k in range(R.shape[0])}
    params = list(model.features.parameters())
    if trainable_decoder:     # decoder.fix: False -- autograd also differentiates the MLP weights (second order included)
        for p in model.decoder.parameters():
            p.requires_grad_(True)
        params += list(model.decoder.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    first = None

    def step():
        opt.zero_grad()
        ld = O.mapping_loss(model, mi, gt, poses, LOSS_CFG["loss_type"], LOSS_CFG["weight_sdf"], LOSS_CFG["weight_eik"],
                            LOSS_CFG["weight_fs"], LOSS_CFG["trunc_dist"], grad_method="autograd", eik_trunc_dist=None)
        total = sum(ld.values())
        total.backward()
        opt.step()
        return total

    for i in range(warmup):
        tot = step()
        if i == 0:
            first = float(tot.detach())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, opt
    torch.cuda.empty_cache()
    return {"value": N_POINTS / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms, "steps": steps,
            "second_order": "reference extension gridsample_grad2 (oracle/_ref)" if mode == "plugin" else "torch gather sampler",
            "first_step_total": first,
            "what": "reference op sequence on cuda: F.grid_sample + aten backward + grad2 plugin + cuBLAS MLP + torch Adam, "
                    "same batch / parameters as the product arm"}


def bench_fd(device, steps=20, warmup=3):
    """The same headline step with the eikonal term as the shipped configs define it (grad_method: finitediff, eps 0.024;
    configs/rgbd/scannet.yaml:48-49): miso_mapping_step_fd (four launches) + Adam, against the reference's GPU op
    sequence for it (7 x [F.grid_sample per level + cuBLAS MLP] forward and backward, torch Adam) on the same batch."""
    from miso_b200 import synth
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import GridTrainer
    from oracle import oracle as O
    cfg = dict(LOSS_CFG, grad_method="finitediff", finite_diff_eps=0.024)
    mi, gt, poses = host_batch(0, 0)
    dmi = {k: v.to(device) for k, v in mi.items()}
    dgt = {k: v.to(device) for k, v in gt.items()}
    net = build_model(device, poses, seed=0)
    tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, net, MisoLossMapping(**cfg), None, device=device)

    def timed(fn):
        first = None
        for i in range(warmup):
            out = fn()
            first = out if first is None else first
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, first

    ms, first = timed(lambda: tr.train_step(dmi, dgt))
    ours_total = float(first[3])
    del tr, net
    torch.cuda.empty_cache()
    shapes = O.level_shapes(synth.SCANNET_SUBMAP_BOUND, 0.5, 5, 2, 4)
    dec = O.make_decoder(8)
    dec.load_state_dict({k.replace("network.", ""): v for k, v in synth.decoder_weights(8).items()})
    model = O.OracleGridNet(synth.SCANNET_SUBMAP_BOUND, initial_features(shapes, 0), dec, second_order=False).to(device)
    R, t = poses
    kf = {k: (R[k].to(device), t[k].to(device)) for k in range(R.shape[0])}
    opt = torch.optim.Adam(list(model.features.parameters()), lr=1e-3)

    def ref_step():
        opt.zero_grad()
        ld = O.mapping_loss(model, dmi, dgt, kf, cfg["loss_type"], cfg["weight_sdf"], cfg["weight_eik"], cfg["weight_fs"],
                            cfg["trunc_dist"], finite_diff_eps=0.024, grad_method="finitediff", eik_trunc_dist=None)
        total = sum(ld.values())
        total.backward()
        opt.step()
        return total.detach()

    ms_ref, first_ref = timed(ref_step)
    ref_total = float(first_ref)
    del model, opt
    torch.cuda.empty_cache()
    return {"what": "headline step with the finite-difference eikonal of the shipped configs (eps 0.024): fused four-launch step "
                    "+ Adam vs the reference's GPU op sequence, same batch and parameters",
            "ms_per_step": ms, "points_per_s": N_POINTS / (ms * 1e-3), "reference_gpu_ms_per_step": ms_ref,
            "speedup_vs_reference_gpu": ms_ref / ms, "first_step_total": ours_total, "reference_first_step_total": ref_total,
            "first_step_rel_err": abs(ours_total - ref_total) / abs(ref_total)}


def bench_trainable_decoder(device, steps=20, warmup=3):
    """The headline step with `decoder.fix: False` (grid_net.py:110,126,346-348): miso_mapping_step (loss terms + grid
    gradients) + miso_mapping_step_wgrad (decoder-parameter gradients incl. the eikonal term's second-order path) + fused
    Adam over grids and decoder, against the reference's GPU op sequence training the same parameters."""
    from miso_b200.loss import MisoLossMapping
    from miso_b200.trainer import GridTrainer
    mi, gt, poses = host_batch(0, 0)
    dmi = {k: v.to(device) for k, v in mi.items()}
    dgt = {k: v.to(device) for k, v in gt.items()}
    net = build_model(device, poses, seed=0)
    for p in net.decoder.parameters():
        p.requires_grad_(True)
    tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, net, MisoLossMapping(**LOSS_CFG), None,
                     device=device)
    first = None
    for _ in range(warmup):
        out = tr.train_step(dmi, dgt)
        first = out if first is None else first
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.train_step(dmi, dgt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ours_total = float(first[3])
    del tr, net
    torch.cuda.empty_cache()
    ref = torch_gpu_arm(device, steps=5, warmup=2, trainable_decoder=True)
    return {"what": "headline step with a TRAINABLE decoder (decoder.fix: False): fused step + decoder-gradient pass + Adam "
                    "vs the reference's GPU op sequence, same batch and parameters",
            "ms_per_step": ms, "points_per_s": N_POINTS / (ms * 1e-3), "reference_gpu_ms_per_step": ref["ms_per_step"],
            "speedup_vs_reference_gpu": ref["ms_per_step"] / ms, "first_step_total": ours_total,
            "reference_first_step_total": ref["first_step_total"],
            "first_step_rel_err": abs(ours_total - ref["first_step_total"]) / abs(ref["first_step_total"])}


def run_torch_gpu(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(0)
    r = torch_gpu_arm(torch.device("cuda", 0), steps=max(1, min(args.steps, 20)), warmup=max(2, min(args.warmup, 3)))
    line = {"impl": "torch-gpu", "metric": "SDF train pts/s (grid+MLP fwd/bwd/eikonal)", "value": r["value"], "unit": "points/s",
            "n_gpus": 1, "steps": r["steps"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "dtype": "f32",
            "data": "synthetic", "config": workload_config(1), "detail": r}
    print(json.dumps(line), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 1))
    cb = cpu_reference_arm(steps=steps, warmup=warm)
    cb.pop("first_step_terms", None)
    line = {"impl": "reference", "metric": "SDF train pts/s (grid+MLP fwd/bwd/eikonal)", "value": cb["value"],
            "unit": "points/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": steps, "warmup": warm,
            "ms_per_step": cb["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(int(os.environ.get("WORLD_SIZE", "1"))),
                           reference_arm="the reference's torch CPU path (oracle port, oracle/oracle.py) on a bounded "
                                         "sample of this workload, rank 0's host cores"),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the alignment half of the metric")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-gpu":
        run_torch_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
