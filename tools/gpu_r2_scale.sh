#!/bin/bash
# round-2 scaling pass on an 8-GPU box: multi-rank parity tests, bench.py at N = 1, 2, 4, 8 (strong
# scaling of the d4 headline, c2/c3 under 'workloads'), config 5 sharded at N = 1, 2, 4, 8
tag=${1:-r02f}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_shadows.py -m gpu -x -q 2>&1 | tail -6
for f in gpurun_out/dist_worker_*.log; do echo "== $f"; grep -E "FAILURES|DIST_GPU_OK" $f | head -3; done
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    def line(tag, w):
        r = w['roofline']
        print(tag, 'value %.3e' % w['value'], 'ms/step %.3f' % w['ms_per_step'],
              'e2e %.3e (%.3f ms)' % (w['e2e']['value'], w['e2e']['ms_per_step']),
              'kernel_ms %.3f' % r['kernel_ms'], 'frac %.3f exec %.3f share %.2f' % (r['frac'], r['executed_frac'], r['kernel_share_of_step']),
              'launches', w['gpu_launches'], 'parity', w['parity'])
    line('HEAD n=%d' % d['n_gpus'], d)
    for k, w in d['workloads'].items():
        line(k, w)
    print('exchange:', d['exchange'], d['clocks'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
n=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  if [ $n -ge $N ]; then
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err || tail -5 $out/bench_n1.err
      timeout 300 python tools/bench_c5_sharded.py --steps 4 > $out/c5_n1.json 2> $out/c5_n1.err || tail -5 $out/c5_n1.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_n$N.json 2> $out/bench_n$N.err || tail -15 $out/bench_n$N.err
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29560+N)) tools/bench_c5_sharded.py --steps 4 > $out/c5_n$N.json 2> $out/c5_n$N.err || tail -15 $out/c5_n$N.err
    fi
    summ $out/bench_n$N.json
    cat $out/c5_n$N.json
  fi
done
