#!/bin/bash
# Times the headline kernel for every library under build_variants/ (scratch tool; run on the GPU box).
# usage: tools/sweep_libs.sh [grid]     env per variant: a file build_variants/<name>.env with VAR=value lines
for so in build_variants/*.so; do
  envf=${so%.so}.env
  ( [ -f $envf ] && export $(cat $envf | xargs); SOLAROSC_LIB=$PWD/$so python bench.py --grid ${1:-96} --steps 2 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$so', 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'grouped %.3e'%d['grouped']['value'], 'grad %.3e'%d['value_and_gradient']['value'], 'frac %.3f'%d['roofline']['frac'])" )
done
