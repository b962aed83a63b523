# round-2 final pass on one B200: full GPU test suite, smoke under compute-sanitizer, bench, ncu launch list + captures
mkdir -p gpurun_out
tag=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; tail -12 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 600 gpurun_out/${tag}_bench_n1.json; grep -n "Error" gpurun_out/${tag}_bench_n1.err | head -3
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/${tag}_bench_reference.json
# launch list of the bench command (cold-cache, serialised per-launch times: the SHARES are what must agree)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_launches.log 2>&1
# alignment level 1: one full capture of align_batch_kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_batch_kernel -s 50 -c 1 -o gpurun_out/${tag}_ncu_align_l1 -f python benchmarks/align_breakdown.py > gpurun_out/${tag}_ncu_align.log 2>&1
# memcheck over every kernel family (smoke) and over a multi-tile mapping step + the new round-2 kernels
timeout 900 compute-sanitizer --launch-timeout 600 --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/${tag}_compute_sanitizer.txt 2>&1; grep -E "ERROR SUMMARY|smoke\]" gpurun_out/${tag}_compute_sanitizer.txt | tail -4
timeout 900 compute-sanitizer --launch-timeout 600 --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_fused.py -q -k "slab or finite_difference or graphed or nan_total or unregistered" >> gpurun_out/${tag}_compute_sanitizer.txt 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${tag}_compute_sanitizer.txt | tail -3
ls -la gpurun_out | grep ${tag}_
