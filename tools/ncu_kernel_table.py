"""Per-launch table from an ncu report: python tools/ncu_kernel_table.py <file.ncu-rep> -> duration, DRAM bytes, issue %, pipes, regs."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "us"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%_active"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
        ("launch__occupancy_limit_registers", "ctas/SM(regs)"), ("launch__occupancy_limit_shared_mem", "ctas/SM(smem)")]
idx = [(hdr.index(k), n, rows[1][hdr.index(k)]) for k, n in want if k in hdr]
for r in rows[2:]:
    print("; ".join(f"{n}={r[i][:70]}{(' ' + u) if u and n not in ('kernel',) else ''}" for i, n, u in idx))
