"""Per-kernel summary of an ncu --set full report (raw page CSV)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"),
        ("sm__cycles_elapsed.max", "cycles"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_insts"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%")]
for d in data:
    print("; ".join(f"{n}={d[idx[k]]}{'' if units[idx[k]] in ('', '%') else ' ' + units[idx[k]]}" for k, n in want if k in idx))
