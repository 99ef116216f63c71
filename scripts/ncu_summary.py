#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + per-instruction hot spots.
usage: python scripts/ncu_summary.py gpurun_out/prof_v3.ncu-rep [out.txt]"""
import csv, subprocess, sys, io

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
        "sm__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
print(f"# {rep}", file=out)
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:92s} {units[i]:16s} {' | '.join(r[i] for r in data)}", file=out)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
first = []
for r in rows[2:]:
    if len(r) < 10:
        if first:
            break
        continue
    first.append(r)

def I(r, k):
    try:
        return int(r[ix[k]])
    except Exception:
        return 0

tot = sum(I(r, "# Samples") for r in first)
execd = sum(I(r, "Instructions Executed") for r in first)
print(f"\n# first launch: {len(first)} SASS lines, {execd} warp instructions, {tot} stall samples", file=out)
items = [(I(r, "# Samples"), I(r, "L1 Wavefronts Shared"), I(r, "L1 Wavefronts Shared Ideal"), I(r, "Instructions Executed"), r[ix["Source"]].strip()) for r in first]
print("# top instructions by stall samples", file=out)
for s, w, wi, ex, txt in sorted(items, reverse=True)[:24]:
    print(f"{s:7d} {100 * s / max(tot, 1):5.1f}%  exec={ex:10d}  {txt}", file=out)
print("# shared-memory wavefronts by instruction", file=out)
for s, w, wi, ex, txt in sorted(items, key=lambda t: -t[1])[:14]:
    if w:
        print(f"wavefronts={w:11d} ideal={wi:11d} exec={ex:10d} per_exec={w / max(ex, 1):5.2f}  {txt}", file=out)
