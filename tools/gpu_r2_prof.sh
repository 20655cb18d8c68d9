#!/bin/bash
# round-2 profiling pass (1 GPU): bench lines, launch list of one bench run (d4 headline + c2 + c3), full ncu
# captures of the control-matrix kernel of every workload, reference arm, c5 local-phase trace
tag=${1:-r02g}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err || tail -5 $out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err || tail -5 $out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
for wl in d4 c3 c2; do
  pat=ctrlmat_static; [ $wl = c2 ] && pat=ctrlmat_dfma
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f -o $out/prof_$wl python bench.py --workload $wl --extra none --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_${wl}_full.log 2>&1
done
echo "== c5 local phase at n_omega = 1250 (what one of 8 ranks does)"
FFB_TRACE=1 timeout 300 python tools/bench_c5_sharded.py --n-omega 1250 --steps 3 2>&1 | grep -E "trace|workload" | tail -8 | tee $out/c5_1250.txt
timeout 300 python tools/time_c3_decay.py 2>&1 | tail -1 | tee $out/c3_decay.txt
ls -la $out
