#!/usr/bin/env python
"""Decode `cuobjdump -sass` output into (address, stall, yield, wr/rd barrier, wait mask, text) rows.

usage: python tools/sass_ctrl.py file.sass [lo_addr hi_addr]   (addresses in hex)
"""
import re
import sys

INS = re.compile(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/")
CTL = re.compile(r"^\s+/\* 0x([0-9a-f]{16}) \*/")


def parse(path):
    rows, pend = [], None
    for line in open(path):
        m = INS.search(line)
        if m:
            pend = (int(m.group(1), 16), m.group(2).strip())
            continue
        m = CTL.match(line)
        if m and pend:
            hi = int(m.group(1), 16)
            rows.append(dict(addr=pend[0], text=pend[1], stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1,
                             wr=(hi >> 46) & 7, rd=(hi >> 49) & 7, wait=(hi >> 52) & 0x3F))
            pend = None
    return rows


if __name__ == "__main__":
    rows = parse(sys.argv[1])
    lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
    for r in rows:
        if lo <= r["addr"] <= hi:
            wr = "-" if r["wr"] == 7 else str(r["wr"])
            rd = "-" if r["rd"] == 7 else str(r["rd"])
            print(f"{r['addr']:05x} st{r['stall']:2d} {'Y' if r['yld'] else ' '} w{wr} r{rd} m{r['wait']:02x}  {r['text']}")
