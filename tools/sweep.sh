#!/bin/bash
# One line per build_variants/lib_*.so: BASELINE metric on a short scan (scratch tool; numbers are for A/B only).
G=${1:-128}
for so in build_variants/lib_*.so; do
  SOLAROSC_LIB=$PWD/$so python bench.py --grid $G --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$so', 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'kernel ms %.2f'%d['roofline']['ms_per_step'], 'grouped %.3e'%d['grouped']['value'], 'frac %.3f'%d['roofline']['frac'], 'parity', d.get('parity_on_sample'))
except Exception as e:
    print('$so', 'FAILED', e)"
done
