mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_full.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","clocks")}, d["e2e"]["value"], d["e2e_compact"]["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["cpu_baseline"]["value"])
for k,v in d["extra"]["align"].items(): print(k, v["iters_per_s"], v["ms_per_iter"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tc2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu.log 2>&1
bash tools/gpu/prof.sh prof_tc2_final MISO_DBG=0
