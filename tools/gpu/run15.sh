mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_full.log
python benchmarks/dense_query.py 2>&1 | tail -1
python benchmarks/scatter_probe.py 2>&1 | tail -1
