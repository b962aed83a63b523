mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_full.log
timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_c.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_c.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["e2e_compact"]["value"], d["e2e_compact"]["h2d_bytes_per_step"], d["e2e_compact"]["ms_per_step"])
PY
