mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
d=json.loads(open(f"gpurun_out/bench_n{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_compact"]["value"])
for k,v in d["extra"]["align"].items(): print(k, v["iters_per_s"], v["ms_per_iter"], v["pairs"], v["pairs_overlapping"], v["samples_per_iter"], v["cuda_graph"])
PY
