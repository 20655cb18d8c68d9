#!/bin/bash
# round-2 check on a box with >= 2 GPUs: GPU tests (incl. the multi-rank parity test), bench at N = 1, 2
tag=${1:-r02a}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    def line(tag, w, clocks=None):
        r = w['roofline']
        print(tag, 'value %.3e' % w['value'], 'ms/step %.3f' % w['ms_per_step'],
              'e2e %.3e (%.3f ms)' % (w['e2e']['value'], w['e2e']['ms_per_step']),
              'kernel_ms %.3f' % r['kernel_ms'], 'frac %.3f exec %.3f share %.2f' % (r['frac'], r['executed_frac'], r['kernel_share_of_step']),
              'launches', w['gpu_launches'], 'parity', w['parity'], clocks or '')
    line('HEAD n=%d' % d['n_gpus'], d, d['clocks'])
    for k, w in d['workloads'].items():
        line(k, w)
    print('exchange:', d['exchange'], '| cpu:', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('kind'))
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err || tail -5 $out/bench_n1.err
summ $out/bench_n1.json
n=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
  if [ $n -ge $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps 10 --warmup 3 > $out/bench_n$N.json 2> $out/bench_n$N.err || tail -15 $out/bench_n$N.err
    summ $out/bench_n$N.json
  fi
done
