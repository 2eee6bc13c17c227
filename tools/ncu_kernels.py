"""Key per-kernel metrics of every launch in an .ncu-rep (raw page)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); h = rows[0]
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__shared_mem_config_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'l1tex__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']
for r in rows[2:]:
    d = dict(zip(h, r))
    print("==", d.get('Kernel Name'))
    for k in KEYS:
        print('  %-72s %s' % (k, d.get(k)))
    for k in h:
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
            v = float(d[k] or 0)
            if v >= 0.05: print('  stall %-66s %.3f' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
