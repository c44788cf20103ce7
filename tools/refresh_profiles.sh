#!/bin/bash
# Turns the gpurun_out/<tag>_* files of tools/profile_round.sh into the tracked summaries under profiles/.
T=${1:-r01e}
cp gpurun_out/${T}_bench_1gpu.json profiles/${R:-r02}_bench_1gpu.json
cp gpurun_out/${T}_bench_reference_arm.json profiles/${R:-r02}_bench_reference_arm.json
cp gpurun_out/${T}_launches.csv profiles/${R:-r02}_launches.csv
python tools/traffic_from_raw.py gpurun_out/${T}_llh_raw.csv profiles/${R:-r02}_llh_kernel_traffic.json > /dev/null
CYC=$(python profiles/raw_summary.py gpurun_out/${T}_llh_raw.csv | grep "smsp__cycles_active.avg" | awk '{print int($2)}')
(python profiles/raw_summary.py gpurun_out/${T}_llh_raw.csv; echo; python profiles/sass_summary.py gpurun_out/${T}_llh_src.csv; echo
 echo "regions of equal execution count: share of sampled warp time, FP64 pipe cycles per sub-partition (DMMA counted 16), pipe utilisation if the region ran alone on all resident warps"
 python tools/ncu_regions.py gpurun_out/${T}_llh_src.csv $CYC 592 4) > profiles/${R:-r02}_llh_kernel_ncu_summary.txt
[ -f gpurun_out/latency.txt ] && cp gpurun_out/latency.txt profiles/${R:-r02}_latency.txt
