#!/bin/bash
# A/B of runtime knobs on one library (scratch tool): tools/envsweep.sh <lib.so> <grid> "VAR=val VAR2=val" ...
so=$1; G=$2; shift 2
for cfg in "$@"; do
  env $cfg SOLAROSC_LIB=$PWD/$so python bench.py --grid $G --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$so [$cfg]', 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'kernel ms %.2f'%d['roofline']['ms_per_step'], 'grouped %.3e'%d['grouped']['value'], 'frac %.3f'%d['roofline']['frac'])
except Exception as e:
    print('$so [$cfg]', 'FAILED', e)"
done
