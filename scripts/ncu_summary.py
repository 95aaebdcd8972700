"""ncu_summary.py -- text summary of an .ncu-rep (raw metrics of interest + per-opcode stall samples from the source page).
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [kernel-index ...] > profiles/...txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = [int(x) for x in sys.argv[2:]]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
kernels = rows[2:]
for n, r in enumerate(kernels):
    if want and n not in want:
        continue
    print(f"===== launch {n}: {r[ix['Kernel Name']]}")
    for k in KEYS:
        if k in ix:
            print(f"{k:75s} {units[ix[k]]:14s} {r[ix[k]]}")
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            try:
                v = float(r[ix[k]])
            except ValueError:
                continue
            if v >= 0.3:
                print(f"{k:75s} {'inst':14s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
all_lines = list(csv.reader(io.StringIO(src)))
starts = [i for i, l in enumerate(all_lines) if l and l[0] == "Kernel Name"] + [len(all_lines)]
for n in range(len(starts) - 1):
    if want and n not in want:
        continue
    lines = all_lines[starts[n]:starts[n + 1]]
    h = lines[1]
    hx = {x: i for i, x in enumerate(h)}
    data = [l for l in lines[2:] if len(l) == len(h)]
    stk = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = 0
    for r in data:
        s = r[hx["Source"]].split()
        if not s:
            continue
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        try:
            smp = int(r[hx["# Samples"]]); ex = int(r[hx["Instructions Executed"]])
        except ValueError:
            continue
        a = agg[op]
        a[0] += smp; a[1] += ex; tot += smp
        for k in stk:
            try:
                a[2][k[6:]] += int(r[hx[k]])
            except ValueError:
                pass
    print(f"----- launch {n}: stall samples by opcode (total {tot})")
    allr = collections.Counter()
    for op, a in sorted(agg.items(), key=lambda x: -x[1][0])[:14]:
        print(f"{op:10s} {a[0]:6d} {100.0 * a[0] / max(tot, 1):5.1f}%  exec={a[1]:10d}  " + ", ".join(f"{k}:{v}" for k, v in a[2].most_common(4)))
    for a in agg.values():
        allr.update(a[2])
    print(dict(allr.most_common(12)))
