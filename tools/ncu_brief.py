"""ncu_brief.py <report.ncu-rep> -- the handful of raw metrics and the hottest SASS of one capture"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) > 0.3:
        print(f"{h:75s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))[2:]
tot = sum(int(r[5]) for r in rows)
print("total warp instr", tot, "sass lines", len(rows))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
for i, r in enumerate(rows):
    c = int(r[5])
    if c > thr * tot:
        print(i, r[1].strip()[:64].ljust(64), c, f"{100*c/tot:.1f}%", "smp", r[4], "thr", r[8])
