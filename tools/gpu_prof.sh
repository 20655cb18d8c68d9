#!/bin/bash
# full ncu capture of the control-matrix kernel for one workload: tools/gpu_prof.sh <workload> <tag>
wl=${1:-d4}; tag=${2:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ctrlmat_main -s 3 -c 1 -o gpurun_out/prof_${wl}_${tag} -f python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${wl}_${tag}.log 2>&1
tail -2 gpurun_out/ncu_${wl}_${tag}.log
