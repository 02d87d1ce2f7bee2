#!/usr/bin/env python
"""Prints, for an .ncu-rep capture, the busiest units, the stall breakdown and a few launch facts (reads ncu --page raw --csv)."""
import csv, io, subprocess, sys
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u, v = rows[0], rows[1], rows[-1]
    d = {a: (b, c) for a, b, c in zip(h, u, v)}
    print("==", path, d.get("Kernel Name", ("", ""))[1][:70])
    top = []
    for a, (b, c) in d.items():
        if "pct_of_peak_sustained_elapsed" in a and ".sum." in a and c not in ("", "0"):
            try: top.append((float(c.replace(",", "")), a))
            except ValueError: pass
    seen = set()
    for val, a in sorted(top, reverse=True):
        if round(val, 2) in seen: continue
        seen.add(round(val, 2))
        print(f"  {val:7.2f} {a}")
        if len(seen) >= 9: break
    st = [(float(c), a.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for a, (b, c) in d.items()
          if "issue_stalled" in a and a.endswith("ratio") and c not in ("", "0")]
    print("  stalls:", ", ".join(f"{a} {val:.2f}" for val, a in sorted(st, reverse=True)[:7]))
    for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"):
        if k in d: print(f"  {k} = {d[k][1]} {d[k][0]}")
