mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_golden.py -m gpu -q > gpurun_out/pytest_align.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_align.log
python benchmarks/align_breakdown.py 2>&1 | tail -1
