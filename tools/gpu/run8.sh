mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_golden.py tests/test_gpu_align.py -m gpu -x -q > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_new.log
python benchmarks/dense_query.py > gpurun_out/dense_query.json 2> gpurun_out/dense_query.err; cat gpurun_out/dense_query.json; tail -3 gpurun_out/dense_query.err
