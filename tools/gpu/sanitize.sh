# compute-sanitizer over (a) the smoke test, (b) a mapping step with several tiles per group, (c) the alignment glue tests
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 compute-sanitizer "$@" > gpurun_out/sanitize_$name.log 2>&1; echo "$name rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_$name.log | tail -1; }
run race_big --tool racecheck --print-limit 20 python tools/gpu/sanitize_big.py 200000
run mem_big --tool memcheck --print-limit 20 python tools/gpu/sanitize_big.py 200000
run mem_align --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_align.py -m gpu -q -k "fused_pose or intersection_counts or iterations_match"
run mem_misc --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "tracker or atlas or compact"
grep -E "terms|passed|failed" gpurun_out/sanitize_*.log | tail -6
