mkdir -p gpurun_out
python benchmarks/scatter_probe.py > gpurun_out/scatter_probe.json 2> gpurun_out/scatter_probe.err; cat gpurun_out/scatter_probe.json; tail -3 gpurun_out/scatter_probe.err
