#!/bin/bash
# One-line summary of a short bench run (scratch tool): value, e2e, per-kernel ms, grouped, FP64 peak.
python bench.py --grid ${1:-96} --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], 'osc ms %.2f'%d['roofline']['ms_per_step'], 'rates ms %.2f'%((d['roofline_rates'] or {}).get('ms_per_step',0)), 'grouped %.3e'%d['grouped']['value'], 'peak %.1f'%d['roofline']['peak'], 'frac %.3f'%d['roofline']['frac'], d['clocks'])"
