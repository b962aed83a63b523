set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_tc2.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_tc2.log
for cfg in "MISO_TC=2" "MISO_TC2_GROUPS=3" "MISO_TC=1"; do
  env $cfg timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["final_loss_terms"])
except Exception as e:
    print("$cfg failed", e); print(open("gpurun_out/bench_$cfg.err").read()[-2000:])
PY
done
