#!/bin/bash
# int8 control-matrix path: parity tests, then timing against the FP64 path
out=gpurun_out/${1:-i8}
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_int8.py -m gpu -x -q -s 2>&1 | tail -25
for v in 0 1; do
  FFB_CTRLMAT_INT8=$v timeout 300 python bench.py --workload d4 --extra c3 --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_i8_$v.json 2> $out/bench_i8_$v.err || tail -5 $out/bench_i8_$v.err
  python - $out/bench_i8_$v.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for tag, w in [('d4', d)] + list(d['workloads'].items()):
    print(tag, 'ms/step %.3f' % w['ms_per_step'], 'e2e %.3f ms' % w['e2e']['ms_per_step'], 'kernel_ms %.3f' % w['roofline']['kernel_ms'], 'parity', w['parity'])
PY
done
