"""Print the handful of `ncu --page raw --csv` metrics the profiles/ summaries quote.  usage: python tools/ncu_keys.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("#", d.get("Kernel Name", ""))
    for k in KEYS:
        if k in d:
            print(f"{k:90s} {d[k]:>22s} {u[k]}")
    st = sorted(((float(v.replace(',', '')), k) for k, v in d.items() if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and v), reverse=True)
    for v, k in st[:12]:
        print(f"{k:90s} {v:22.0f}")
