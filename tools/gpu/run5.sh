mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_tc2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tc2.log
bash tools/gpu/ablate.sh "MISO_DBG=0" "MISO_PAIR=0" "MISO_DBG=3" "MISO_TC2_GROUPS=3 MISO_DBG=0" "MISO_TC2_GROUPS=3 MISO_PAIR=0"
