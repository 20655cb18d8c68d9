#!/bin/bash
# Measurement pass on the GPU box (one GPU): tools/gpu_measure.sh <tag>
#   bench lines (C2 default, C3, d4, reference arm), launch lists, full ncu captures of the three
#   dominant kernels (C2: ctrlmat_dfma, d4: ctrlmat_static, C5: from_atomic_dmma), sequencing timings.
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
python bench.py --steps 20 --warmup 5 > $out/bench_c2.json 2> $out/bench_c2.err; tail -2 $out/bench_c2.err
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_c3.json 2> $out/bench_c3.err
python bench.py --workload d4 --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_d4.json 2> $out/bench_d4.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference_c2.json 2> $out/bench_ref.err
python tools/bench_sequencing.py --cpu > $out/sequencing.jsonl 2> $out/sequencing.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_d4.csv python bench.py --workload d4 --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_d4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_c5.csv python tools/run_c5.py 10000 1 > $out/ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctrlmat_dfma -s 3 -c 1 -f -o $out/prof_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_c2_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctrlmat_static -s 3 -c 1 -f -o $out/prof_d4 python bench.py --workload d4 --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_d4_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:from_atomic_dmma -s 0 -c 1 -f -o $out/prof_c5_atomic python tools/run_c5.py 10000 1 > $out/ncu_c5_full.log 2>&1
ls -la $out
