#!/bin/bash
# Times the oscillation kernel for each launch-bound variant built into build_variants/ (scratch tool).
for so in build_variants/lib_*.so; do
  SOLAROSC_LIB=$PWD/$so python bench.py --grid 96 --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$so', 'value %.3e'%d['value'], 'osc ms %.2f'%d['roofline']['ms_per_step'], 'rates ms %.2f'%d['roofline_rates']['ms_per_step'], 'grouped %.3e'%d['grouped']['value'], 'peak %.1f'%d['roofline']['peak'])"
done
