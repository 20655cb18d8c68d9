#!/bin/bash
# round-2 pass c (1 GPU): all GPU tests, RB per-call loop, c5 concatenate with trace
tag=${1:-r02c}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== RB per-call loop"
timeout 300 python tools/profile_rb_loop.py 2>&1 | head -30 | tee $out/rb_loop.txt
echo "== c5 concatenate (single GPU, trace)"
FFB_TRACE=1 timeout 300 python tools/run_c5.py 10000 4 2>&1 | grep -v "pulse pipeline\|control matrix:" | tail -12 | tee $out/c5_concatenate.txt
echo "== sequencing bench"
timeout 600 python tools/bench_sequencing.py > $out/sequencing.jsonl 2> $out/sequencing.err; cat $out/sequencing.jsonl
