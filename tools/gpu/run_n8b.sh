mkdir -p gpurun_out
for nb in 1 0; do
MISO_NUMA_BIND=$nb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 60 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_n8_numa$nb.json 2> gpurun_out/bench_n8_numa$nb.err; echo "rc=$?"
python - $nb <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/bench_n8_numa{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("numa_bind", sys.argv[1], d["host_numa_node_rank0"], d["value"], d["e2e"]["value"], d["e2e_compact"]["value"])
PY
done
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket" | head
