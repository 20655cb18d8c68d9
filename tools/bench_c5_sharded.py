#!/usr/bin/env python
"""BASELINE config 5 with omega sharded over the GPUs of one node: the 4-qubit QFT (d = 16, 256-element
GGM basis, 18 noise operators) concatenated from its gate pulses, n_omega frequencies in total (STRONG
scaling: the grid is fixed, every rank takes 1/N of it).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/bench_c5_sharded.py [--n-omega 10000] [--steps 5]

A step = cache the 9 gate control matrices on the rank's block, ff.concatenate them there (local work,
no exchange), then gather the filter function on every rank (51.8 MB at n_omega = 1e4).  Wall clock per
phase, max over ranks; rank 0 prints one JSON line.  Run without torchrun for N = 1."""
import argparse
import json
import os
import sys
import time

# one JSON line on stdout: NCCL prints its version there at start-up, send everything else to stderr
sys.stdout.flush()
RESULT_OUT = os.fdopen(os.dup(1), 'w')
os.dup2(2, 1)

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n-omega', type=int, default=10_000)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import filter_functions_b200 as ff
    from filter_functions_b200 import distributed as ffd
    world = int(os.environ.get('WORLD_SIZE', 1))
    if world > 1:
        ffd.init_process_group('nccl')
    rank = dist.get_rank() if world > 1 else 0
    omega = np.logspace(-2, 2, args.n_omega)
    pulses = workloads.build_qft_pulses(ff, 4)
    o0, o1 = ffd.owned_frequencies(len(omega), rank, world)
    start, stop = ffd.frequency_shard(len(omega), rank, world)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step():
        """-> (pulse, gathered F, seconds of local work, seconds of the gather)"""
        for p in pulses:
            p.cleanup('frequency dependent')
        barrier()
        t0 = time.perf_counter()
        new, _ = ffd.concatenate(pulses, omega, gather=False)
        F_local = new.get_filter_function(omega[o0:o1])
        barrier()
        t1 = time.perf_counter()
        F = ffd.allgather_frequency_axis(F_local, len(omega), halo=False) if world > 1 else F_local
        barrier()
        t2 = time.perf_counter()
        return new, F, t1 - t0, t2 - t1

    for _ in range(args.warmup):
        new, F, _, _ = step()
    local, gather = [], []
    for _ in range(args.steps):
        new, F, tl, tg = step()
        local.append(tl)
        gather.append(tg)
    total = [a + b for a, b in zip(local, gather)]
    t = torch.tensor([min(total), float(np.median(total)), float(np.median(local)),
                      float(np.median(gather))], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # parity of the gathered result: Hermitian; equal to the from-scratch filter function of the
    # concatenated pulse on a few frequencies; and (rank 0) equal to the single-GPU concatenation on the
    # whole grid
    herm = float(np.abs(F - F.conj().transpose(1, 0, 2)).max()/np.abs(F).max())
    pick = np.linspace(0, args.n_omega - 1, 7).astype(int)
    bare = ff.concatenate(pulses, calc_filter_function=False)
    F_s = bare.get_filter_function(omega[pick])
    err = float(np.abs(F[..., pick] - F_s).max()/np.abs(F_s).max())
    single = None
    if rank == 0 and world > 1:
        for p in pulses:
            p.cleanup('frequency dependent')
            p.cache_control_matrix(omega)
        F_1 = ff.concatenate(pulses, omega=omega).get_filter_function(omega)
        single = float(np.abs(F - F_1).max()/np.abs(F_1).max())
    if world > 1:
        dist.barrier()
    if rank == 0:
        peers = ffd.peer_group() if world > 1 else None
        print(json.dumps({
            'workload': f'c5: QFT from 9 gate pulses (d=16, 256 basis elements, 18 noise operators), '
                        f'n_omega={args.n_omega} sharded over {world} GPU(s), F gathered on every rank',
            'n_gpus': world, 'scaling': 'strong', 's_per_step_best': t[0].item(),
            's_per_step_median': t[1].item(), 's_local_median': t[2].item(),
            's_gather_median': t[3].item(), 'gathered_bytes': int(F.nbytes), 'steps': args.steps,
            'exchange': ('none' if world == 1 else 'peer windows over NVLink (ffb_allgather_columns)'
                         if peers is not None else 'NCCL all_gather_into_tensor'),
            'gathered_F_hermiticity': herm, 'gathered_vs_scratch_max_rel_diff': err,
            'sharded_vs_single_gpu_max_rel_diff': single}), file=RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
