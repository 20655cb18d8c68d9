#!/bin/bash
# session 2e: Gram kernels (guard-free stages), whole GPU suite, default bench with the d4_int8 line
out=gpurun_out/${1:-s2e}
mkdir -p $out
(time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) 2>&1 | tee $out/pytest.log
timeout 300 python tools/time_ff_kernel.py > $out/ff_gram.jsonl 2> $out/ff_gram.err || tail -3 $out/ff_gram.err
cat $out/ff_gram.jsonl
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err || tail -5 $out/bench_default.err
python - $out/bench_default.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for tag, w in [('d4', d)] + list(d['workloads'].items()):
    print(tag, 'ms/step %.3f' % w['ms_per_step'], 'e2e %.3f ms' % w['e2e']['ms_per_step'], 'kernel_ms %.3f' % w['roofline']['kernel_ms'], 'parity', w['parity'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_gram -s 2 -c 1 -f -o $out/prof_ff_gram python tools/time_ff_kernel.py > $out/ncu_gram_full.log 2>&1
ls -la $out
