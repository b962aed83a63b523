# usage: bash tools/gpu/ablate.sh "<env settings>" ...   -> one line per configuration
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-extras > gpurun_out/abl.json 2> gpurun_out/abl.err
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/abl.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:40s} step {d['ms_per_step']:.4f} ms  kernel {d['roofline']['kernel_ms']:.4f} ms  frac {d['roofline']['frac']:.3f}  loss {d['final_loss_terms'][3]:.6f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/abl.err").read()[-1500:])
PY
done
