# decoder-gradient pass: smoke, the BASELINE-size test, the bench extra
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -x -k "trainable_decoder" 2>&1 | tail -3
timeout 600 python - <<'PY' 2>&1 | tail -5
import json, torch, bench
r = bench.bench_trainable_decoder(torch.device("cuda:0"))
print(json.dumps(r))
open("gpurun_out/r02_trainable_decoder.json", "w").write(json.dumps(r))
PY
