# quick re-check after a library change: fused + golden GPU tests, bench N=1 (record), launch list
mkdir -p gpurun_out
tag=${1:-r02}
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_golden.py tests/test_gpu_align.py -q 2>&1 | tail -3
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 400 gpurun_out/${tag}_bench_n1.json; grep -n "Error" gpurun_out/${tag}_bench_n1.err | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_launches.log 2>&1
