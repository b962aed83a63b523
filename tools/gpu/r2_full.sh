# round-2 full pass on one B200: GPU tests, bench, interp sweep (with the reference's grad2 kernel), ncu captures
mkdir -p gpurun_out
tag=${1:-r02a}
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; tail -15 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -c 3000 gpurun_out/${tag}_bench_n1.json; tail -5 gpurun_out/${tag}_bench_n1.err
timeout 600 python benchmarks/interp_sweep.py > gpurun_out/${tag}_interp_sweep.csv 2> gpurun_out/${tag}_interp_sweep.err; grep -c . gpurun_out/${tag}_interp_sweep.csv; tail -3 gpurun_out/${tag}_interp_sweep.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mapping_step_tc2 -s 3 -c 1 -o gpurun_out/${tag}_ncu_scannet -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_scannet.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mapping_step_tc2 -s 3 -c 1 -o gpurun_out/${tag}_ncu_ncd -f python benchmarks/ncd_point_sharded.py > gpurun_out/${tag}_ncu_ncd.log 2>&1
ls -la gpurun_out | grep ${tag}
