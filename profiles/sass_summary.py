#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: executed warp instructions and stall samples per SASS opcode.

usage: python profiles/sass_summary.py <source_page.csv> [top_n]
"""
import collections
import csv
import re
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    h = next(k for k, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    i_s, i_i, i_w = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp = collections.Counter(), collections.Counter()
    tot = tots = 0.0
    for r in rows[h + 1:]:
        if len(r) <= i_i:
            continue
        try:
            n, w = float(r[i_i] or 0), float(r[i_w] or 0)
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i_s])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += n
        samp[op] += w
        tot += n
        tots += w
    print(f"total warp instructions {tot:.4g}, stall samples {tots:.4g}")
    for op, n in ops.most_common(top):
        print(f"{op:<10s} {n / tot * 100:5.1f}% inst  {samp[op] / max(tots, 1) * 100:5.1f}% samples")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
