"""Summarise an .ncu-rep: key utilisation metrics per captured launch (reads `ncu --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    print("-" * 100)
    for w in want:
        if w in idx:
            print("%-68s %s %s" % (w, r[idx[w]], units[idx[w]]))
    st = []
    for h, i in idx.items():
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and r[i] not in ("", "n/a"):
            st.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
    print("stalls/issue:", ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:7]))
