#!/usr/bin/env python
"""Summarise an .ncu-rep (no GPU needed): key raw metrics + per-source-line instruction / stall shares.
usage: python tools/ncu_summary.py REPORT.ncu-rep [kernel-regex] > profiles/xyz.txt"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]


def run(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
for r in raw[2:]:
    name = r[hdr.index("Kernel Name")]
    import re
    if not re.search(kre, name):
        continue
    print("== kernel:", name[:100])
    for w in WANT:
        if w in hdr:
            print(f"  {w:78s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
    break

src = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}",
                                      "--launch-count", "1"))))
cur, lines, sass = None, [], []
for r in src:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name"):
        continue
    try:
        if r[0].strip().isdigit():
            lines.append((cur, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0)))
        elif r[2].startswith("0x"):
            sass.append((r[3].strip(), int(r[6] or 0), int(r[7] or 0)))
    except ValueError:
        pass
ti = sum(x[2] for x in sass) or 1
ts = sum(x[1] for x in sass) or 1
print(f"\n== SASS totals: {ti} warp-instructions, {ts} stall samples")
ops = Counter()
for s_, sm, ie in sass:
    p = s_.split()
    if not p:
        continue
    op = p[1] if p[0].startswith("@") and len(p) > 1 else p[0]
    ops[op.split(".")[0]] += ie
print("== opcode mix (share of warp-instructions)")
for op, v in ops.most_common(24):
    print(f"  {op:12s} {v:11d} {100 * v / ti:5.1f}%")
print("== hottest SASS by stall samples")
for s_, sm, ie in sorted(sass, key=lambda x: -x[1])[:18]:
    print(f"  {100 * sm / ts:5.1f}%  inst={ie:9d}  {s_[:90]}")
print("== source lines (inst share incl. inlined callees / stall-sample share)")
tli = sum(x[4] for x in lines) or 1
tls = sum(x[3] for x in lines) or 1
for f, l, s_, sm, ie in lines:
    if ie > tli * 0.008 or sm > tls * 0.01:
        print(f"  {f[:16]:16s}:{l:4d} inst {100 * ie / tli:5.1f}% stall {100 * sm / tls:5.1f}% | {s_[:96]}")
