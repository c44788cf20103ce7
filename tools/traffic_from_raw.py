#!/usr/bin/env python
"""DRAM traffic of the profiled llh_kernel launch from an `ncu --page raw --csv` export, as the small JSON that
bench.py reads for `roofline.traffic` (bytes per parameter point: the profiled launch and the timed launches differ
in size, the per-point figure carries over).

usage: python tools/traffic_from_raw.py gpurun_out/<tag>_llh_raw.csv profiles/<tag>_llh_kernel_traffic.json
"""
import csv
import json
import sys


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    rd, wr, grid = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), int(val("launch__grid_size"))
    out = {"kernel": r[hdr.index("Kernel Name")][:40], "dram_bytes_read": rd, "dram_bytes_write": wr, "points_in_launch": grid,
           "dram_bytes_per_point": (rd + wr) / grid, "algorithmic_bytes_per_point": 19 * 8 + 12,
           "source": "ncu --set full --clock-control none, one llh_kernel launch of bench.py --grid 64 (%s)" % src}
    json.dump(out, open(dst, "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main(*sys.argv[1:3])
