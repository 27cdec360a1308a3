#!/usr/bin/env python
"""`.ncu-rep` (ncu --set full capture brought back in gpurun_out/) -> markdown table of the metrics the
roofline discussion uses + the warp-state sample breakdown.   python tools/ncu_summary.py rep [title]"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    rec = dict(zip(hdr, vals))
    un = dict(zip(hdr, units))
    print("### %s — `%s` (launch ID %s)\n" % (title, rec.get("Kernel Name", "?"), rec.get("ID", "?")))
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEEP:
        if k in rec and rec[k] != "":
            print("| `%s` | %s | %s |" % (k, rec[k], un.get(k, "")))
    stalls = [(k, float(rec[k].replace(",", ""))) for k in hdr
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and rec.get(k, "") not in ("", "n/a")]
    tot = sum(v for _, v in stalls) or 1.0
    stalls.sort(key=lambda kv: -kv[1])
    print("\nWarp-state samples (%d): %s\n" % (tot, ", ".join("%s %.1f %%" % (k[len("smsp__pcsamp_warps_issue_stalled_"):], 100 * v / tot)
                                                        for k, v in stalls[:10])))
