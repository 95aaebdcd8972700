"""prints the headline metrics of the first kernel in an ncu report:  python scripts/ncu_metrics.py report.ncu-rep"""
import csv, subprocess, sys, io
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread ', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct', 'lts__t_bytes.sum ', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled', 'sm__throughput.avg.pct', 'smsp__inst_executed.sum ', 'l1tex__throughput.avg.pct', 'sm__cycles_elapsed.max']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if any(w in h + ' ' for w in WANT) and 'Triage' not in h:
            if 'issue_stalled' in h and float(v or 0) < 0.3:
                continue
            print(f"{h:92s} {u:10s} {v}")
    print('-' * 60)
