"""Summarise an .ncu-rep: duration, pipe use, stall reasons per issue (measurement tooling)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "")[:70])
    stalls = []
    for h, v in d.items():
        if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            stalls.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    print("  stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    for k in ("gpu__time_duration.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
              "smsp__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"):
        if k in d: print("  ", k, d[k], units[hdr.index(k)])
