#!/bin/bash
# correctness + quick numbers: GPU tests, then the three bench workloads
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for wl in c2 c3 d4; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err || tail -5 gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$wl.json'))
    r=d['roofline']
    print('$wl', 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e (%.2f ms)'%(d['e2e']['value'], d['e2e']['ms_per_step']), 'kernel_ms %.3f'%r['kernel_ms'], 'frac %.3f exec_frac %.3f share %.2f'%(r['frac'], r['executed_frac'], r['kernel_share_of_step']), 'launches', d['gpu_launches'], 'parity', d['parity_device_vs_api'], d['clocks'])
except Exception as e:
    print('$wl failed', e)
PY
done
