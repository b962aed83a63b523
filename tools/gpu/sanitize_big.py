"""racecheck/memcheck target: the two-thread mapping kernel with several tiles per group (cross-tile hazards)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import SMALL_BOUND, make_pair
from miso_b200 import synth
from miso_b200.loss import MisoLossMapping
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
net, _, _ = make_pair(device="cuda:0")
mi, gt, (R, t) = synth.rgbd_batch(N, num_kf=4, bound=SMALL_BOUND, seed=1, wall_margin=0.3)
for k in range(R.shape[0]):
    net.set_initial_kf_pose(k, R[k], t[k], kf_key=f"KF{k}")
net.unlock_feature(); net.lock_pose()
L = MisoLossMapping(loss_type="L1", weight_sdf=1.0, weight_eik=0.5, weight_fs=0.1, trunc_dist=0.15, grad_method="autograd", eik_trunc_dist=0.1)
out = L.step_into_grads(net, {k: v.cuda() for k, v in mi.items()}, {k: v.cuda() for k, v in gt.items()})
torch.cuda.synchronize()
print("terms", out.tolist())
