"""Summarise gpurun_out ncu artefacts into small text files under profiles/ (tracked)."""
import collections, csv, subprocess, sys

def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        agg.setdefault(r[ki].split('(')[0], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, 'w') as f:
        f.write(f"# per-launch device time from `ncu --metrics gpu__time_duration.sum --clock-control none` ({path})\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':42s} {'launches':>8s} {'mean_us':>10s} {'share_%':>8s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k[:42]:42s} {len(v):8d} {sum(v)/len(v)/1e3:10.1f} {100*sum(v)/tot:8.1f}\n")

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sass__inst_executed_local',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__average_warps_issue_stalled', 'launch__shared_mem_per_block']

def full(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write(f"# selected metrics from `ncu --set full --clock-control none` ({rep})\n")
        for row in rows[2:]:
            f.write(f"## kernel: {row[h.index('Kernel Name')]}  grid={row[h.index('Grid Size')]} block={row[h.index('Block Size')]}\n")
            for i, n in enumerate(h):
                if any(w in n for w in WANT) and 'per_second' not in n and 'pct_of_peak_sustained_elapsed' not in n or n in ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):
                    f.write(f"{n:85s} {units[i]:>16s} {row[i]}\n")

if __name__ == '__main__':
    kind, src, dst = sys.argv[1:4]
    (launches if kind == 'launches' else full)(src, dst)
