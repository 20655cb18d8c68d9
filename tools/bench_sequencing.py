#!/usr/bin/env python
"""Timing of BASELINE.json configs 4 and 5 (parity-test cases of the main bench, measured here for
DESIGN.md): randomized-benchmarking sequences from cached Cliffords and the 4-qubit QFT concatenated
from cached gate pulses.  One JSON line per measurement.

    python tools/bench_sequencing.py [--rb-sequences 1000] [--qft-omega 10000] [--cpu]

``--cpu`` also times the oracle port of the reference algorithm (bounded sample) on the host cores.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402


def best_of(fn, n=3):
    out, best = None, 1e300
    for _ in range(n):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rb-sequences', type=int, default=1000)
    ap.add_argument('--rb-loop', type=int, default=100, help='sequences timed through ff.concatenate')
    ap.add_argument('--qft-omega', type=int, default=10_000)
    ap.add_argument('--cpu', action='store_true')
    args = ap.parse_args()
    import __graft_entry__ as entry
    entry.build()
    import filter_functions_b200 as ff

    # ---- config 4 -----------------------------------------------------------------------------------
    omega = workloads.rb_omega()
    S = workloads.rb_spectrum(omega)
    cliffords = workloads.build_cliffords(ff, omega)
    rows = workloads.rb_sequences(args.rb_sequences)
    ff.concatenate_many(cliffords, rows[:8], spectrum=S)          # warm-up
    t_batch, batch = best_of(lambda: ff.concatenate_many(cliffords, rows, spectrum=S,
                                                         calc_control_matrix=False,
                                                         calc_filter_function=False))
    t_batch_full, _ = best_of(lambda: ff.concatenate_many(cliffords, rows, spectrum=S))

    def loop():
        return [ff.infidelity(ff.concatenate([cliffords[k] for k in row]), S, omega)
                for row in rows[:args.rb_loop]]
    loop()
    t_loop, infids = best_of(loop, 2)
    err = float(np.abs(np.array(infids) - batch.infidelities[:args.rb_loop]).max()
                / np.abs(batch.infidelities).max())
    print(json.dumps({
        'workload': f'c4: {len(rows)} RB sequences of {rows.shape[1]} Cliffords, n_omega=301',
        'concatenate_many_infidelity_only': {'s': t_batch, 'sequences_per_s': len(rows)/t_batch},
        'concatenate_many_all_arrays': {'s': t_batch_full, 'sequences_per_s': len(rows)/t_batch_full},
        'concatenate_loop': {'s_per_sequence': t_loop/args.rb_loop,
                             'sequences_per_s': args.rb_loop/t_loop},
        'loop_vs_batch_max_rel_diff': err}), flush=True)

    # ---- config 5 -----------------------------------------------------------------------------------
    omega = np.logspace(-2, 2, args.qft_omega)
    pulses = workloads.build_qft_pulses(ff, 4)

    def cache_gates():
        for p in pulses:
            p.cleanup('frequency dependent')
            p.cache_control_matrix(omega)
    cache_gates()
    t_gates, _ = best_of(cache_gates, 2)
    ff.concatenate(pulses, omega=omega)
    t_concat, qft = best_of(lambda: ff.concatenate(pulses, omega=omega), 2)

    def scratch():
        bare = ff.concatenate(pulses, calc_filter_function=False)
        return bare.get_filter_function(omega)
    scratch()
    t_scratch, F_scratch = best_of(scratch, 2)
    F = qft.get_filter_function(omega)
    print(json.dumps({
        'workload': f'c5: 4-qubit QFT (d=16, GGM basis 256, 13 segments, 18 noise operators), '
                    f'n_omega={args.qft_omega}',
        'cache_gate_control_matrices_s': t_gates, 'concatenate_s': t_concat,
        'from_scratch_filter_function_s': t_scratch,
        'concat_vs_scratch_max_rel_diff': float(np.abs(F - F_scratch).max()/np.abs(F).max())}),
        flush=True)

    if args.cpu:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import ff_oracle as oracle
        # RB: one sequence through the oracle's from_atomic + filter function + integral
        lib_B = np.array([c.get_control_matrix(workloads.rb_omega()) for c in cliffords])
        lib_ph = np.array([c.get_total_phases(workloads.rb_omega()) for c in cliffords])
        lib_L = np.array([c.total_propagator_liouville for c in cliffords])
        om = workloads.rb_omega()

        def cpu_rb(row):
            phases = lib_ph[row[:-1]].cumprod(axis=0)
            Q = np.empty((len(row) - 1, 4, 4))
            Q[0] = lib_L[row[0]]
            for i in range(1, len(row) - 1):
                Q[i] = lib_L[row[i]] @ Q[i - 1]
            B = oracle.control_matrix_from_atomic(phases, lib_B[row], Q)
            return oracle.infidelity_from_filter_function(oracle.filter_function(B), S, om, 2)
        t_cpu, _ = best_of(lambda: [cpu_rb(r) for r in rows[:50]], 2)
        # QFT: from-scratch control matrix of the 13-segment pulse on a bounded frequency sample
        n_cpu = 200
        bare = ff.concatenate(pulses, calc_filter_function=False)
        H = oracle.hamiltonian_from_coeffs(bare.c_opers, bare.c_coeffs)
        ev, V, Q = oracle.diagonalize(H, bare.dt)
        t0 = time.perf_counter()
        oracle.control_matrix_from_scratch(ev, V, Q, omega[:n_cpu], np.asarray(bare.basis),
                                           bare.n_opers, bare.n_coeffs, bare.dt)
        t_qft = (time.perf_counter() - t0)*len(omega)/n_cpu
        print(json.dumps({
            'cpu_port': {'rb_numeric_kernel_s_per_sequence': t_cpu/50,
                         'rb_note': 'numeric part only (no PulseSequence bookkeeping); the '
                                    'reference spends 21 ms per ff.concatenate call (SURVEY 6)',
                         'qft_from_scratch_s_extrapolated': t_qft,
                         'qft_note': f'{n_cpu} of {len(omega)} frequencies, scaled linearly',
                         'cores': os.cpu_count()}}), flush=True)


if __name__ == '__main__':
    main()
