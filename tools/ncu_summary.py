#!/usr/bin/env python
"""Print the headline metrics + top stall reasons of an .ncu-rep (helper for profiles/)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'lts__t_sectors_op_red.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
print("metric,unit,value")
for i, n in enumerate(h):
    if n in keep:
        print(f"{n},{u[i]},{v[i]}")
st = [(float(v[i]), n) for i, n in enumerate(h) if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio')]
for val, n in sorted(st, reverse=True)[:8]:
    print(f"{n},ratio,{val:.3f}")

# --traffic-key KEY: record dram bytes per launch of this capture in profiles/ncu_traffic.json (read by bench.py)
if "--traffic-key" in sys.argv:
    import json, os
    key = sys.argv[sys.argv.index("--traffic-key") + 1]
    rd = float(v[h.index('dram__bytes_read.sum')]); wr = float(v[h.index('dram__bytes_write.sum')])
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale.get(u[h.index('dram__bytes_read.sum')], 1); wr *= scale.get(u[h.index('dram__bytes_write.sum')], 1)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = {"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
              "kernel": v[h.index('Kernel Name')] if 'Kernel Name' in h else None,
              "duration_us_under_ncu": float(v[h.index('gpu__time_duration.sum')]) if 'gpu__time_duration.sum' in h else None,
              "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", ".csv") + " (ncu --set full, one launch)"}
    json.dump(d, open(path, "w"), indent=1)
