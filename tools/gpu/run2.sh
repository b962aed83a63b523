bash tools/gpu/ablate.sh "MISO_DBG=0" "MISO_DBG=1" "MISO_DBG=2" "MISO_DBG=3" "MISO_DBG=4" "MISO_DBG=7" \
  "MISO_TC2_GROUPS=3 MISO_DBG=0" "MISO_TC2_GROUPS=3 MISO_DBG=1" "MISO_TC2_GROUPS=3 MISO_DBG=2" "MISO_TC2_GROUPS=3 MISO_DBG=3" "MISO_TC2_GROUPS=3 MISO_DBG=7" \
  "MISO_TC=1 MISO_DBG=0" "MISO_TC=1 MISO_DBG=1" "MISO_TC=1 MISO_DBG=2" "MISO_TC=1 MISO_DBG=3"
ncu --set full --clock-control none --import-source on -k regex:mapping_step_tc2 -s 3 -c 1 -o gpurun_out/prof_tc2_g4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_g4.log 2>&1
MISO_TC2_GROUPS=3 ncu --set full --clock-control none --import-source on -k regex:mapping_step_tc2 -s 3 -c 1 -o gpurun_out/prof_tc2_g3 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_g3.log 2>&1
ls -la gpurun_out
