#!/bin/bash
# A/B: an environment switch of the library (name=value pairs given as arguments), config 2 bench line
for setting in "$@"; do
  export $setting
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$setting', 'step %.4f ms'%d['ms_per_step'], 'kernel %.4f ms'%r['kernel_ms'], 'e2e %.3f ms'%d['e2e']['ms_per_step'], 'exec_frac %.3f'%r['executed_frac'], 'parity', d['parity_device_vs_api'])"
done
