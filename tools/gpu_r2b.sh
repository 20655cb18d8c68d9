#!/bin/bash
# round-2 pass b (2 GPUs): all GPU tests, shadow consumers (c3 decay, c5 concatenate), c5 sharded,
# RB per-call loop profile, compute-sanitizer memcheck / racecheck on reduced-size tests
tag=${1:-r02b}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for f in gpurun_out/dist_worker_*.log; do echo "== $f"; grep -E "FAILURES|DIST_GPU_OK|Error" $f | head -5; done
echo "== c3 decay amplitudes from the cached control matrix"
timeout 300 python tools/time_c3_decay.py 2>&1 | tail -2 | tee $out/c3_decay.txt
echo "== c5 concatenate (single GPU)"
timeout 300 python tools/run_c5.py 10000 4 2>&1 | tail -4 | tee $out/c5_concatenate.txt
echo "== c5 sharded"
timeout 300 python tools/bench_c5_sharded.py --steps 4 > $out/c5_n1.json 2> $out/c5_n1.err || tail -5 $out/c5_n1.err
cat $out/c5_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/bench_c5_sharded.py --steps 4 > $out/c5_n2.json 2> $out/c5_n2.err || tail -15 $out/c5_n2.err
cat $out/c5_n2.json
echo "== RB per-call loop"
timeout 300 python tools/profile_rb_loop.py 2>&1 | head -45 | tee $out/rb_loop.txt
echo "== compute-sanitizer memcheck"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $out/memcheck.log python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_numeric.py -m gpu -x -q -k "pipeline or control_matrix_from_scratch or infidelity_integral or identity_basis or special_frequencies or short_pulses" 2>&1 | tail -3
tail -5 $out/memcheck.log
echo "== compute-sanitizer racecheck"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $out/racecheck.log python -m pytest tests/test_gpu_numeric.py -m gpu -x -q -k "control_matrix_from_scratch or infidelity_integral or diagonalize" 2>&1 | tail -3
tail -5 $out/racecheck.log
