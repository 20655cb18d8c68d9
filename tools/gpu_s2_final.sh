#!/bin/bash
# round-2 final single-GPU pass: compute-sanitizer over the kernels new in this session, launch list of a
# bench run, bench line + reference arm
tag=${1:-r02i}
out=gpurun_out/$tag
mkdir -p $out
S=$out/sanitizer.txt
echo "# compute-sanitizer over reduced-size GPU tests (tools/gpu_s2_final.sh, pass $tag)" > $S
echo "## memcheck: filter-function tests (row-pair, Gram 4x4 / 6x3 / panel-pair kernels), int8 control-matrix tests (small sizes), pulse-route bit-identity, long-pulse intermediates" >> $S
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_numeric.py tests/test_gpu_int8.py tests/test_gpu_pulse.py -m gpu -x -q -k "filter_function or int8_control_matrix_against_oracle or identical_bits or long_pulse" 2>&1 | grep -E "=========|passed|failed" | tail -8 >> $S
echo "## racecheck: Gram kernels (cp.async ring + block barrier + shuffle), int8 kernel (mbarrier ring, TMEM)" >> $S
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_numeric.py tests/test_gpu_int8.py -m gpu -x -q -k "filter_function_gram or (int8_control_matrix_against_oracle and 64-6-64)" 2>&1 | grep -E "=========|passed|failed" | tail -8 >> $S
cat $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
python tools/summarize_launches.py $out/launches_bench.csv > $out/launches_bench.txt 2>&1; head -40 $out/launches_bench.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err || tail -5 $out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err || tail -5 $out/bench_reference.err
python - $out/bench_n1.json $out/bench_reference.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
for tag, w in [('d4', d)] + list(d['workloads'].items()):
    print(tag, 'value %.3e' % w['value'], 'ms/step %.3f' % w['ms_per_step'], 'e2e %.3f ms' % w['e2e']['ms_per_step'], 'kernel_ms %.3f' % w['roofline']['kernel_ms'], 'parity', w['parity'])
print('cpu_baseline', d['cpu_baseline'])
r = json.load(open(sys.argv[2]))
print('reference arm', r['value'], r['ms_per_step'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
PY
FFB_CTRLMAT_INT8=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctrlmat_i8_kernel -s 3 -c 1 -f -o $out/prof_i8_d4 python bench.py --workload d4 --extra none --no-int8 --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_i8_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_gram -s 2 -c 1 -f -o $out/prof_ff_gram python tools/time_ff_kernel.py > $out/ncu_gram_full.log 2>&1
timeout 300 python tools/time_ff_kernel.py > $out/ff_kernel.jsonl 2>/dev/null
timeout 300 python tools/diag_propagator_error.py d4 c2 > $out/propagator_error.txt 2>&1
timeout 300 python tools/bench_sequencing.py > $out/sequencing.jsonl 2> $out/sequencing.err || tail -3 $out/sequencing.err
tail -3 $out/sequencing.jsonl | cut -c1-400
ls -la $out
