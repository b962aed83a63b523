# one full ncu capture of the decoder-gradient kernel at the headline batch size (2^20 samples, ScanNet grid)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mapping_wgrad_kernel -s 2 -c 1 -f -o gpurun_out/r02_ncu_wgrad python - > gpurun_out/r02_ncu_wgrad.log 2>&1 <<'PY'
import torch, bench
from miso_b200.loss import MisoLossMapping
from miso_b200.trainer import GridTrainer
dev = torch.device("cuda", 0)
mi, gt, poses = bench.host_batch(0, 0)
dmi = {k: v.to(dev) for k, v in mi.items()}
dgt = {k: v.to(dev) for k, v in gt.items()}
net = bench.build_model(dev, poses, seed=0)
for p in net.decoder.parameters():
    p.requires_grad_(True)
tr = GridTrainer({"learning_rate": 1e-3, "grid_training_mode": "joint"}, net, MisoLossMapping(**bench.LOSS_CFG), None, device=dev)
for _ in range(4):
    tr.train_step(dmi, dgt)
torch.cuda.synchronize()
PY
tail -3 gpurun_out/r02_ncu_wgrad.log; ls -la gpurun_out/r02_ncu_wgrad.ncu-rep
