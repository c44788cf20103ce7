#!/usr/bin/env python
"""For every backward branch in a cuobjdump SASS listing: body length, FP64 instruction count, sum of the
encoded stall counts (a lower bound of one warp's cycles per iteration) and the FP64 pipe time at one
instruction per two cycles (DMMA m8n8k4: 16).  A loop is pipe-bound on an SM sub-partition with W resident
warps when stalls <= W * pipe.

usage: python tools/sass_loops.py file.sass [min_body]
"""
import re
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from sass_ctrl import parse  # noqa: E402


def main():
    rows = parse(sys.argv[1])
    min_body = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    by_addr = {r["addr"]: i for i, r in enumerate(rows)}
    for i, r in enumerate(rows):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:P\d,\s*)?0x([0-9a-f]+)", r["text"])
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt > r["addr"] or tgt not in by_addr:
            continue
        body = rows[by_addr[tgt]: i + 1]
        if len(body) < min_body:
            continue
        n64 = sum(1 for b in body if re.match(r"(@!?P\d\s+)?D(FMA|MUL|ADD|SETP|MNMX)", b["text"]))
        nmma = sum(1 for b in body if "DMMA" in b["text"])
        stalls = sum(max(b["stall"], 1) for b in body)
        pipe = 2 * n64 + 16 * nmma
        nlds = sum(1 for b in body if re.search(r"\bLDS", b["text"]))
        nldg = sum(1 for b in body if re.search(r"\bLD(G|\.E)", b["text"]))
        print(f"loop {tgt:05x}-{r['addr']:05x}: {len(body):4d} instr, {n64:4d} FP64, {nmma:3d} DMMA, {nlds:3d} LDS, {nldg:2d} LDG, "
              f"stalls {stalls:5d}, pipe {pipe:5d}, stalls/pipe {stalls / max(pipe, 1):.2f}")


if __name__ == "__main__":
    main()
