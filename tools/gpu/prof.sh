# usage: bash tools/gpu/prof.sh <name> [env...]   -> gpurun_out/<name>.ncu-rep
mkdir -p gpurun_out
name=$1; shift
env "$@" ncu --set full --clock-control none --import-source on -k regex:mapping_step_tc -s 3 -c 1 -o gpurun_out/$name -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
