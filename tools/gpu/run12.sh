mkdir -p gpurun_out
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_a1.json 2> gpurun_out/bench_a1.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_a1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_a1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
for k,v in d["extra"]["align"].items(): print(k, v)
PY
