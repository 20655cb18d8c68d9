"""Worker of tests/test_distributed_gloo.py: runs under torch.distributed.run with the gloo backend.

Exercises the frequency-sharded path of filter_functions_b200.distributed (shard -> local integral ->
all-reduce; shard -> local filter function -> all-gather) on CPU tensors.  There is no GPU here, so the
per-rank computations the product runs on the rank's GPU (numeric.infidelity,
PulseSequence.get_filter_function, the numerical tail of concatenate) are replaced by the oracle by
patching those names in this worker process -- the product signatures carry no test hooks.  The
multi-rank parity test of the real CUDA path is tests/test_gpu_distributed.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)

import torch.distributed as dist  # noqa: E402

import ff_oracle as oracle  # noqa: E402
import filter_functions_b200 as ff  # noqa: E402
from filter_functions_b200 import distributed as ffd  # noqa: E402
from helpers import rand_pulse_sequence  # noqa: E402


def oracle_filter_function(pulse, omega):
    H = oracle.hamiltonian_from_coeffs(pulse.c_opers, pulse.c_coeffs)
    ev, V, Q = oracle.diagonalize(H, pulse.dt)
    B = oracle.control_matrix_from_scratch(ev, V, Q, omega, np.asarray(pulse.basis), pulse.n_opers,
                                           pulse.n_coeffs, pulse.dt)
    return oracle.filter_function(B)


def oracle_infidelity(pulse, spectrum, omega, n_oper_identifiers=None):
    idx = ff.util.get_indices_from_identifiers(pulse.n_oper_identifiers, n_oper_identifiers)
    return oracle.infidelity_from_filter_function(oracle_filter_function(pulse, omega), spectrum,
                                                  omega, pulse.d, idx)


def patch_device_calls():
    """Route the three device computations of the sharded path to the oracle (CPU-only worker)."""
    real_concatenate = ff.pulse_sequence.concatenate
    ffd.numeric.infidelity = oracle_infidelity
    ff.PulseSequence.get_filter_function = (
        lambda self, omega, *args, **kwargs: oracle_filter_function(self, np.asarray(omega)))
    ffd.pulse_sequence.concatenate = (
        lambda pulses, omega=None, calc_filter_function=None:
        real_concatenate(pulses, calc_filter_function=False))       # host bookkeeping only


def main():
    ffd.init_process_group('gloo')
    patch_device_calls()
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(99)          # same pulse on every rank (operands are replicated)
    pulse = rand_pulse_sequence(ff, rng, 3, 8, 2, 3)
    failures = []
    for n_omega in (1, 2, 3, 64, 129):
        omega = np.geomspace(0.05, 20, n_omega) if n_omega > 1 else np.array([0.3])
        S1 = 1e-2/omega
        S2 = np.array([S1*(k + 1) for k in range(3)])
        S3 = np.einsum('a,b,o->abo', [1, 2, 3], [1, 2, 3], S1) + 0j
        for S in (S1, S2, S3):
            got = ffd.infidelity(pulse, S, omega)
            want = oracle_infidelity(pulse, S, omega)
            if got.shape != want.shape or np.abs(got - want).max() > 1e-13*max(1.0, np.abs(want).max()):
                failures.append(('infidelity', n_omega, S.ndim))
        ids = list(pulse.n_oper_identifiers[[2, 0]])
        got = ffd.infidelity(pulse, S2[:2], omega, n_oper_identifiers=ids)
        want = oracle_infidelity(pulse, S2[:2], omega, ids)
        if np.abs(got - want).max() > 1e-13*max(1.0, np.abs(want).max()):
            failures.append(('infidelity ids', n_omega))
        F = ffd.filter_function(pulse, omega)
        F_want = oracle_filter_function(pulse, omega)
        if F.shape != F_want.shape or np.abs(F - F_want).max() > 1e-13*np.abs(F_want).max():
            failures.append(('filter_function', n_omega))
    # sharded concatenation: every rank concatenates on its own frequency block, F is all-gathered
    pieces = [pulse[0:3], pulse[3:4], pulse[4:8]]

    for n_omega in (1, 2, 5, 64, 131):
        omega = np.geomspace(0.05, 20, n_omega) if n_omega > 1 else np.array([0.3])
        joined, F = ffd.concatenate(pieces, omega)
        F_want = oracle_filter_function(pulse, omega)
        if F.shape != F_want.shape or np.abs(F - F_want).max() > 1e-12*np.abs(F_want).max():
            failures.append(('concatenate', n_omega))
        if abs(joined.tau - pulse.tau) > 1e-12:
            failures.append(('concatenate tau', n_omega))
    total = ffd.allreduce_sum(np.array([float(rank + 1)]))
    if total[0] != world*(world + 1)/2:
        failures.append(('allreduce', total))
    dist.barrier()
    if failures:
        print(f'rank {rank}: FAILURES {failures}', flush=True)
        sys.exit(1)
    if rank == 0:
        print('DIST_OK', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
