"""One rank of tests/test_gpu_multi.py (launched by torch.distributed.run): slab-sharded fit vs the single-GPU trainer."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from helpers import SMALL_BOUND, make_pair  # noqa: E402
from miso_b200 import synth  # noqa: E402
from miso_b200.loss import MisoLossMapping  # noqa: E402
from miso_b200.sharded_fit import SlabShardedFit  # noqa: E402
from miso_b200.trainer import GridTrainer  # noqa: E402


def log(rank, msg):
    print(f"[rank {rank}] {msg}", flush=True)


def main():
    halo, out_dir = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    mi, gt, (R, t) = synth.rgbd_batch(20000, num_kf=4, bound=SMALL_BOUND, seed=11, wall_margin=0.3)
    dmi, dgt = {k: v.to(dev) for k, v in mi.items()}, {k: v.to(dev) for k, v in gt.items()}
    nets = []
    for _ in range(2):
        net, _, _ = make_pair(device=dev)
        for k in range(R.shape[0]):
            net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
        net.unlock_feature()
        net.lock_pose()
        nets.append(net)

    def mk():
        return MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15,
                               grad_method="autograd", eik_trunc_dist=0.1)
    tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, nets[0], mk(), None, device=dev)
    fit = SlabShardedFit(nets[1], mk(), lr=1e-3, halo=halo)
    assert fit.p2p == (halo == "p2p")
    log(rank, f"calibrated {fit.calibrate(dmi)} axis {fit.axis}")
    steps, terms_ref, terms = 7, [], []
    for _ in range(steps):
        terms_ref.append(tr.train_step(dmi, dgt).clone())
    for i in range(2):
        terms.append(fit.step(dmi, dgt).clone())                 # eager, no prefetch
        torch.cuda.synchronize()
        log(rank, f"eager step {i} done")
    replay = fit.graphed_step(dmi, dgt, prefetch_same=True)       # + 1 eager step with prefetch
    log(rank, "captured")
    terms.append(None)
    for _ in range(steps - 3):
        terms.append(replay().clone())                            # two alternating graphs
    torch.cuda.synchronize()
    log(rank, "replays done")
    fit.gather_model()
    fit.restore_layout()
    torch.cuda.synchronize()
    err_p = [float((a - b).norm() / b.norm()) for a, b in zip(nets[1].level_tensors(), nets[0].level_tensors())]
    err_t = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(terms, terms_ref) if a is not None)
    log(rank, f"err_p {err_p} err_t {err_t}")
    torch.save({"err_p": err_p, "err_t": err_t, "own": int(fit._bufs["count"].item()), "slab": [fit.zb, fit.ze]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    # captured graphs hold NCCL work and the peer mappings: tearing the process group down under them can block, and
    # nothing is left to flush -- leave without the interpreter's teardown
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
