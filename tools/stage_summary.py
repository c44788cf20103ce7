#!/usr/bin/env python
"""profiles/r01_stage_kernels_ncu.txt from the raw page of the tools/stage_kernels_profile.py capture.

usage: python tools/stage_summary.py gpurun_out/r01_stage_raw.csv profiles/r01_stage_kernels_ncu.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]


def col(n):
    return hdr.index(n)


def num(r, k):
    v = r[col(k)].replace(",", "")
    return float(v) if v else float("nan")


def tobytes(r, k):
    return num(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[col(k)], 1)


def toms(r, k):
    return num(r, k) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "usecond": 1e-3, "msecond": 1, "nsecond": 1e-6}.get(units[col(k)], 1)


out = []
for r in rows[2:]:
    name = r[col("Kernel Name")].replace("solarosc::", "").replace("solarosc_setup::", "")[:44]
    ms = toms(r, "gpu__time_duration.sum")
    rd, wr = tobytes(r, "dram__bytes_read.sum"), tobytes(r, "dram__bytes_write.sum")
    fp = num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
    out.append(f"{name:46s} grid {r[col('launch__grid_size')]:>7s} regs {r[col('launch__registers_per_thread')]:>4s} {ms * 1e3:9.1f} us  "
               f"fp64 pipe {fp:5.1f} %  dram rd {rd / 1e6:8.2f} MB wr {wr / 1e6:8.2f} MB = {(rd + wr) / ms / 1e6:7.1f} GB/s")
text = ("ncu --set full --clock-control none of tools/stage_kernels_profile.py: the stage-separated kernels, launches in program order\n"
        "(16 384-point likelihood batch with SOLAROSC_NO_FUSED=1; 1 024-sample derived block; 64 x 80 400-term Poisson vectors;\n"
        "16 384-point inclusive + angular batch).  Cold-cache, serialised launches: compare shares and fractions, not absolutes.\n\n"
        + "\n".join(out) + "\n")
open(sys.argv[2], "w").write(text)
print(text)
