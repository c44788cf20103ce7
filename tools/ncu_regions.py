#!/usr/bin/env python
"""Per-region FP64 pipe utilisation from an `ncu --page source --csv` export (SASS view).

A region is a run of instructions with (nearly) equal execution count.  For each: share of stall samples
(= share of warp time), FP64 pipe time of its instructions (2 cycles per warp instruction per sub-partition,
16 for DMMA) as a share of all sampled warp time / resident warps, i.e. how busy the pipe would be if only
this region ran.

usage: python tools/ncu_regions.py source.csv total_cycles n_smsp warps_per_smsp
"""
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    h = next(k for k, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    i_s, i_i, i_w = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    out = []
    for r in rows[h + 1:]:
        if len(r) <= i_i:
            continue
        try:
            out.append((int(float(r[i_i] or 0)), r[i_s], int(float(r[i_w] or 0))))
        except ValueError:
            pass
    return out


def main():
    data = load(sys.argv[1])
    cycles, n_smsp, wps = float(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
    tots = sum(d[2] for d in data)
    segs, start = [], 0
    for i in range(1, len(data) + 1):
        if i == len(data) or abs(data[i][0] - data[start][0]) > 0.02 * max(data[start][0], 1):
            segs.append((start, i - 1))
            start = i
    print("region        exec/inst  len  time%  pipe-clk/SMSP  wall-clk   pipe-util")
    acc = []
    for a, b in segs:
        body = data[a:b + 1]
        smp = sum(d[2] for d in body)
        pipe = 0.0
        for ex, txt, _ in body:
            if re.search(r"\bDMMA", txt):
                pipe += 16.0 * ex
            elif re.search(r"\bD(FMA|MUL|ADD|SETP)", txt):
                pipe += 2.0 * ex
        pipe /= n_smsp
        wall = smp / tots * cycles   # if all resident warps of the sub-partition sat in this region together
        acc.append((a, b, body[0][0], smp / tots, pipe, wall))
    for a, b, ex, sh, pipe, wall in acc:
        if sh > 0.004 or pipe > 0.004 * cycles:
            print(f"{a:5d}-{b:5d} {ex:11.3e} {b - a + 1:4d} {100 * sh:5.1f}% {pipe:12.0f} {wall:10.0f}   {pipe / max(wall, 1):.2f}")
    print("total pipe", sum(x[4] for x in acc), "of", cycles)


if __name__ == "__main__":
    main()
