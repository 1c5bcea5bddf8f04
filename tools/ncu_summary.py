#!/usr/bin/env python
"""Condense an ncu --set full report into the per-kernel summary CSV kept under profiles/:
  python tools/ncu_summary.py gpurun_out/<tag>_prof_eqplane.ncu-rep profiles/<tag>_ncu_summary.csv"""
import csv
import io
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
ki = hdr.index("Kernel Name")
names = ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
         "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
         "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
         "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
names += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
names += ["l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
cols = rows[2:]
out = [["metric", "unit"] + [r[ki].split("(")[0].replace("void ", "") for r in cols]]
for n in names:
    if n in hdr:
        i = hdr.index(n)
        out.append([n, rows[1][i]] + [r[i] for r in cols])
i1, i2, i3 = [hdr.index("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k) for k in ("dadd", "dmul", "dfma")]
out.append(["derived: executed FP64 flop per cycle / 18944 (148 SM x 64 FMA x 2)", "frac"] +
           ["%.4f" % ((float(r[i1]) + float(r[i2]) + 2 * float(r[i3])) / 18944) for r in cols])
csv.writer(open(out_path, "w")).writerows(out)
for o in out:
    print(o[0][:72], o[2:])
