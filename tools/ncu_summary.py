"""Prints the handful of ncu raw metrics we track from a .ncu-rep (run where ncu is installed, no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op", "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_alu", "gpu__dram_throughput", "lts__throughput.avg.pct",
        "smsp__inst_executed.avg.per_cycle_active", "sm__cycles_active.avg", "launch__grid_size", "sm__inst_executed_pipe_fmaheavy", "local"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("==", name)
    for h, v in zip(hdr, r):
        if any(k in h for k in keys) and "pcsamp" not in h and ".min" not in h and ".max" not in h and "per_second" not in h:
            print(f"  {h} = {v}")
