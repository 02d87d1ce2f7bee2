#!/usr/bin/env python
"""Summarises ncu captures (gpurun_out/*.ncu-rep) into profiles/: per capture a small CSV of the metrics the design
argues with (duration, DRAM bytes, instruction counts, issue utilisation, occupancy) — and updates profiles/traffic.json
for the headline kernel, stamped with the hash of the kernel source the capture was taken from.
usage: python tools/ncu_summary.py <name>=<path.ncu-rep> ..."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[-1]


def main():
    res = {}
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        hdr, units, vals = raw(path)
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        with open(os.path.join(ROOT, "profiles", f"{name}_ncu.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "unit", "value"])
            w.writerow(["Kernel Name", "", d.get("Kernel Name", ("", ""))[1]])
            for k in KEEP:
                if k in d:
                    w.writerow([k, d[k][0], d[k][1]])
        res[name] = d

        def num(k):
            u, v = d[k]
            v = float(v.replace(",", ""))
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}.get(u, 1.0)
            return v * scale
        print(f"{name}: {d['Kernel Name'][1][:60]}  {num('gpu__time_duration.sum')*1e6:.1f} us  read {num('dram__bytes_read.sum')/1e6:.1f} MB  write {num('dram__bytes_write.sum')/1e6:.1f} MB  "
              f"inst {num('smsp__inst_executed.sum')/1e6:.1f} M  issue {d['smsp__issue_active.avg.pct_of_peak_sustained_active'][1]} %")
        if name == "r2_cfg2":
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            tj = json.load(open(tp))
            rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
            tj["fused_resize_cfg2_bytes_per_launch"] = int(rd + wr)
            tj["fused_resize_cfg2_source_sha16"] = hashlib.sha256(open(os.path.join(ROOT, "kornia-rs_b200", "csrc", "resize_fused.cu"), "rb").read()).hexdigest()[:16]
            tj["fused_resize_cfg2_capture"] = "profiles/r2_cfg2_ncu.csv (ncu --set full, 64 frames 3840x2160 -> 1280x720, one launch)"
            tj["fused_resize_cfg2"].update({"dram_read_B": int(rd), "dram_write_B": int(wr), "capture": tj["fused_resize_cfg2_capture"]})
            json.dump(tj, open(tp, "w"), indent=1)


if __name__ == "__main__":
    main()
