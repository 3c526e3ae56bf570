#!/usr/bin/env python
"""Basic-block view of an ncu capture (made with --set full --import-source on): groups the SASS of the captured
kernel into runs of instructions with the same execution count and prints, per run, the executed warp-instructions,
the average number of active threads, shared-memory wavefronts, stall samples and the opcode mix.  This is the view
DESIGN.md §8's per-block numbers come from: it shows which code regions a warp walks at how many active lanes.

  python tools/ncu_blocks.py gpurun_out/prof_agg_r1_v6.ncu-rep [min_share_percent]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[start - 1][1][:150] if start else "")
    hdr, body = rows[start], [r for r in rows[start + 1:] if len(r) > 10]
    col = {n: hdr.index(n) for n in ("Instructions Executed", "Avg. Threads Executed", "L1 Wavefronts Shared", "# Samples")}

    def num(r, n):
        try:
            return float(r[col[n]])
        except ValueError:
            return 0.0
    total = sum(num(r, "Instructions Executed") for r in body)
    print(f"{total / 1e6:.1f} M warp-instructions, {len(body)} SASS instructions")
    runs, cur = [], None
    for k, r in enumerate(body):
        key = (num(r, "Instructions Executed"), num(r, "Avg. Threads Executed"))
        toks = r[1].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        if cur and cur["key"] == key:
            cur["n"] += 1
        else:
            cur = dict(key=key, at=k, n=1, wf=0.0, samples=0.0, ops=collections.Counter())
            runs.append(cur)
        cur["wf"] += num(r, "L1 Wavefronts Shared"); cur["samples"] += num(r, "# Samples"); cur["ops"][op] += 1
    tot_samples = sum(x["samples"] for x in runs) or 1.0
    print(" idx  len   exec/inst  thr    warp-inst  share   smem-wf  stall%  opcodes")
    for x in runs:
        ex, thr = x["key"]
        share = 100.0 * ex * x["n"] / total
        if share < min_share and 100.0 * x["samples"] / tot_samples < min_share:
            continue
        print(f"{x['at']:5d} {x['n']:4d} {ex / 1e6:9.2f}M {thr:5.1f} {ex * x['n'] / 1e6:9.1f}M {share:5.1f}% {x['wf'] / 1e6:8.1f}M "
              f"{100.0 * x['samples'] / tot_samples:5.1f}%  {dict(x['ops'].most_common(5))}")


if __name__ == "__main__":
    main()
