# last pass of round 2 on one B200: full GPU suite, N=1 bench record, memcheck of smoke (every kernel family incl. the decoder-gradient pass)
mkdir -p gpurun_out
tag=${1:-r02}
timeout 900 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${tag}_pytest.log 2>&1; tail -10 gpurun_out/${tag}_pytest.log
timeout 700 python bench.py --steps 200 --warmup 10 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 500 gpurun_out/${tag}_bench_n1.json; grep -n "Error" gpurun_out/${tag}_bench_n1.err | head -3
timeout 600 compute-sanitizer --launch-timeout 600 --tool memcheck --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/${tag}_compute_sanitizer.txt 2>&1; grep -E "ERROR SUMMARY|smoke\]" gpurun_out/${tag}_compute_sanitizer.txt | tail -4
