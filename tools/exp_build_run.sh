#!/bin/bash
# usage: tools/exp_build_run.sh "<extra nvcc flags>" <workload>...   (timing experiments; results may be wrong)
flags="$1"; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo $flags -shared -Xcompiler -fPIC -o filter_functions_b200/csrc/libffb200.so filter_functions_b200/csrc/*.cu || exit 1
touch filter_functions_b200/csrc/libffb200.so
for wl in "$@"; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl', 'kernel_ms %.3f exec_frac %.3f step %.3f'%(r['kernel_ms'], r['executed_frac'], d['ms_per_step']))"
done
