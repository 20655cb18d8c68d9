#!/bin/bash
# multi-GPU check: tools/gpu_multi.sh N  (weak-scaling bench lines for C2 and d4 at N GPUs)
N=${1:-2}
mkdir -p gpurun_out
for wl in c2 d4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
  tail -3 gpurun_out/bench_${wl}_n$N.err; cat gpurun_out/bench_${wl}_n$N.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 | tail -2
