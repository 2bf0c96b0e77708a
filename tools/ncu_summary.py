#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) for the kernels it holds. Usage: ncu_summary.py file.ncu-rep [topN]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
for r in rows[2:]:
    print("=====", r[idx["Kernel Name"]][:90])
    for w in WANT:
        if w in idx:
            print("  %-86s %s %s" % (w, r[idx[w]], units[idx[w]]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
sections = []; cur = None
for row in csv.reader(io.StringIO(src)):
    if not row: continue
    if row[0] == "File Path": cur = {"file": row[1], "rows": [], "hdr": None}; sections.append(cur); continue
    if row[0] == "Function Name": cur["func"] = row[1]; continue
    if row[0] == "Line No": cur["hdr"] = row; continue
    if cur and cur["hdr"]: cur["rows"].append(row)
by = collections.OrderedDict()
for s in sections:
    h = s["hdr"]; iL = h.index("Line No"); iI = h.index("Instructions Executed"); iS = h.index("# Samples")
    fn = s["func"].split("(")[0][-40:]
    for r in s["rows"]:
        try:
            ok = r[iL].strip().isdigit() and float(r[iI] or 0) > 0
        except (ValueError, IndexError):
            continue  # a source line with embedded quotes / commas (inline asm) that the CSV writer split
        if ok:
            by.setdefault(fn, []).append((float(r[iI]), s["file"].split("/")[-1], int(r[iL]), r[1].strip()[:105], float(r[iS] or 0) if r[iS] not in ('-', '') else 0.0))
for fn, rows_ in by.items():
    tot = sum(r[0] for r in rows_); ts = sum(r[4] for r in rows_) or 1
    print("==== %s: %.3e warp-instructions attributed (inlined lines are double counted)" % (fn, tot))
    for inst, f, ln, text, samp in sorted(rows_, reverse=True)[:topn]:
        print("%6.2f%% inst %6.2f%% samples  %-18s L%-4d %s" % (100 * inst / tot, 100 * samp / ts, f[:18], ln, text))
