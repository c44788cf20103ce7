#!/usr/bin/env python
"""Split an `ncu --page source --csv` SASS export into regions of equal execution count (basic-block groups)
and print each region's share of executed warp instructions; optionally dump one region's SASS.

usage: python profiles/sass_regions.py <source_page.csv> [first_instr last_instr]
"""
import csv
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
    if len(sys.argv) > 3:
        for k in range(int(sys.argv[2]), int(sys.argv[3]) + 1):
            print(f"{k:5d} {data[k][0]:12d} {data[k][2]:6d}  {data[k][1]}")
        return
    tot = sum(d[0] for d in data)
    tots = sum(d[2] for d in data)
    prev, start, acc, accs = None, 0, 0, 0
    segs = []
    for i, d in enumerate(data):
        if prev is None or abs(d[0] - prev) > 0.02 * max(prev, 1):
            if prev is not None:
                segs.append((start, i - 1, prev, acc, accs))
            start, acc, accs = i, 0, 0
        acc += d[0]
        accs += d[2]
        prev = d[0]
    segs.append((start, len(data) - 1, prev, acc, accs))
    for s in segs:
        if s[3] > 0.01 * tot:
            print(f"instr {s[0]:5d}-{s[1]:5d}  exec/inst {s[2]:.3e}  inst share {100 * s[3] / tot:5.1f}%  "
                  f"sample share {100 * s[4] / max(tots, 1):5.1f}%  len {s[1] - s[0] + 1}")


if __name__ == "__main__":
    main()
