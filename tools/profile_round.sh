#!/bin/bash
# Round-end evidence on the GPU box: bench (no profiler), reference arm, ncu launch list, one full capture of
# llh_kernel and of the other fused kernels.  usage: tools/profile_round.sh <tag> [quick]   (quick: only the llh capture)
T=${1:-r02}
if [ "$2" != "quick" ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench_1gpu.json 2> gpurun_out/${T}_bench.err
  timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${T}_bench_reference_arm.json 2>> gpurun_out/${T}_bench.err
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --grid 64 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_l.log 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:'pipe_kernel|gen_llh_kernel|grad_kernel|adjoint_kernel' -s 4 -c 8 -f -o gpurun_out/${T}_others \
    python bench.py --grid 64 --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${T}_ncu_o.log 2>&1
  ncu -i gpurun_out/${T}_others.ncu-rep --page raw --csv > gpurun_out/${T}_others_raw.csv
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:llh_kernel -s 3 -c 1 -f -o gpurun_out/${T}_llh_kernel \
  python bench.py --grid 64 --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/${T}_ncu_s.log 2>&1
ncu -i gpurun_out/${T}_llh_kernel.ncu-rep --page raw --csv > gpurun_out/${T}_llh_raw.csv
ncu -i gpurun_out/${T}_llh_kernel.ncu-rep --page source --csv > gpurun_out/${T}_llh_src.csv
tail -3 gpurun_out/${T}_ncu_s.log | cut -c1-300
