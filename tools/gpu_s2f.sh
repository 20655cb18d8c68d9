#!/bin/bash
# session 2f: Gram kernel with 16-element stages; the reference's own test-suite against this package
out=gpurun_out/${1:-s2f}
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_numeric.py -m gpu -x -q -k "filter_function" 2>&1 | tail -3
timeout 300 python tools/time_ff_kernel.py > $out/ff_gram.jsonl 2> $out/ff_gram.err || tail -3 $out/ff_gram.err
head -4 $out/ff_gram.jsonl
timeout 1500 python tools/run_reference_tests.py --out $out/reference_tests.json 2>&1 | tail -150
