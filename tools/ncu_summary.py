"""Summarise .ncu-rep files (one kernel launch each) into a small CSV: usage ncu_summary.py out.csv rep1 rep2 ..."""
import csv, subprocess, sys
METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__cycles_elapsed.max",
           "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
           "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_active.avg.per_cycle_active",
           "launch__shared_mem_per_block_dynamic"]
out = csv.writer(open(sys.argv[1], "w"))
out.writerow(["report", "kernel"] + METRICS)
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    d = {n: f"{v[i]} {u[i]}".strip() for i, n in enumerate(h)}
    out.writerow([rep.split("/")[-1], d.get("Kernel Name", "")] + [d.get(m, "") for m in METRICS])
