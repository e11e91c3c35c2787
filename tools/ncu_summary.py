"""Key numbers of .ncu-rep captures (usage: ncu_summary.py a.ncu-rep b.ncu-rep ...) and raw-page CSV export."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    if "--export" in sys.argv:
        continue
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
    print("==", rep, name[:80])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print("   %-72s %s %s" % (k, v[i], u[i]))
