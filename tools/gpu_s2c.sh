#!/bin/bash
# session 2c: Gram kernel v2 (tests, timing, profile); int8 kernel contention experiments
out=gpurun_out/${1:-s2c}
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_numeric.py -m gpu -x -q -k "filter_function" 2>&1 | tail -5
timeout 300 python tools/time_ff_kernel.py > $out/ff_gram.jsonl 2> $out/ff_gram.err || tail -3 $out/ff_gram.err
cat $out/ff_gram.jsonl
for dbg in 4 8 5; do
  FFB_CTRLMAT_INT8=1 FFB_I8_DEBUG=$dbg timeout 300 python bench.py --workload d4 --extra none --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_i8_dbg$dbg.json 2> $out/bench_i8_dbg$dbg.err || tail -5 $out/bench_i8_dbg$dbg.err
  python - $out/bench_i8_dbg$dbg.json $dbg <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print('i8 debug', sys.argv[2], 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_gram -s 2 -c 1 -f -o $out/prof_ff_gram python tools/time_ff_kernel.py > $out/ncu_gram_full.log 2>&1
ls -la $out
