#!/bin/bash
# experiment: number of frequency blocks of the fused pipeline (download of block i overlaps block i+1)
for wl in c2 c3 d4; do
for nb in 1 2 3 4 8; do
  export FFB_PIPELINE_BLOCKS=$nb
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl blocks $nb', 'step %.4f ms'%d['ms_per_step'], 'kernel %.4f ms'%r['kernel_ms'], 'e2e %.3f ms'%d['e2e']['ms_per_step'], 'parity', d['parity_device_vs_api'])"
done
done
unset FFB_PIPELINE_BLOCKS
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
