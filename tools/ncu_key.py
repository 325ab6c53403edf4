"""Key metrics of the first kernel in an .ncu-rep.  usage: python tools/ncu_key.py x.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines())); h = rows[0]; v = rows[2]
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'derived__memory_l1_wavefronts_shared_excessive', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'smsp__warps_eligible.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
          'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']:
    print(f"{k:75s} {v[h.index(k)] if k in h else None}")
out = []
for i, k in enumerate(h):
    if 'stalled' in k and 'ratio' in k and 'not_issued' not in k:
        try: out.append((round(float(v[i]), 2), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
        except ValueError: pass
print("stalls per issue:", sorted(out, reverse=True)[:9])
