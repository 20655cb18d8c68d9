#!/bin/bash
# measurement pass on the GPU box: GPU tests, smoke, bench lines, launch list, full ncu captures
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
python bench.py --workload d4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d4.json 2> gpurun_out/bench_d4.err; tail -3 gpurun_out/bench_d4.err; cat gpurun_out/bench_d4.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_d4.csv python bench.py --workload d4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctrlmat_main -s 3 -c 1 -f -o gpurun_out/prof_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctrlmat_main -s 3 -c 1 -f -o gpurun_out/prof_d4 python bench.py --workload d4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d4_full.log 2>&1
ls -la gpurun_out
