"""Worker of tests/test_gpu_distributed.py: one rank per GPU under torch.distributed.run (NCCL).

The REAL sharded CUDA path -- every rank computes its frequency block on its own GPU through the C ABI,
the partial integrals are summed in the epilogue of the infidelity kernel over NVLink peer memory (or by
NCCL with FFB_PEER=0), F(omega) blocks are gathered through the peer windows -- checked on the GLOBAL
grid against the oracle and the reference's golden fixtures.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ff_oracle as oracle  # noqa: E402
import filter_functions_b200 as ff  # noqa: E402
import workloads  # noqa: E402
from filter_functions_b200 import distributed as ffd  # noqa: E402
from helpers import nerr, rand_pulse_sequence  # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
TOL = 1e-10


def oracle_filter_function(pulse, omega):
    H = oracle.hamiltonian_from_coeffs(pulse.c_opers, pulse.c_coeffs)
    ev, V, Q = oracle.diagonalize(H, pulse.dt)
    B = oracle.control_matrix_from_scratch(ev, V, Q, omega, np.asarray(pulse.basis), pulse.n_opers,
                                           pulse.n_coeffs, pulse.dt)
    return oracle.filter_function(B)


def oracle_infidelity(pulse, spectrum, omega, ids=None):
    idx = ff.util.get_indices_from_identifiers(pulse.n_oper_identifiers, ids)
    return oracle.infidelity_from_filter_function(oracle_filter_function(pulse, omega), spectrum,
                                                  omega, pulse.d, idx)


def main():
    ffd.init_process_group('nccl')
    rank, world = dist.get_rank(), dist.get_world_size()
    failures = []

    def check(name, got, want, tol=TOL):
        got, want = np.ascontiguousarray(got), np.ascontiguousarray(want)
        if got.shape != want.shape or not np.isfinite(got.view(float)).all() or nerr(got, want) > tol:
            failures.append((name, got.shape, want.shape,
                             nerr(got, want) if got.shape == want.shape else None))

    peers = ffd.peer_group()
    want_peers = os.environ.get('FFB_PEER', '1') != '0'
    if want_peers and peers is None:
        failures.append('peer group could not be created (CUDA IPC / P2P)')
    if not want_peers and peers is not None:
        failures.append('FFB_PEER=0 ignored')

    total = ffd.allreduce_sum(np.arange(5.0) + rank)
    check('allreduce', total, world*np.arange(5.0) + world*(world - 1)/2, 1e-15)

    rng = np.random.default_rng(99)          # same pulse on every rank (operands are replicated)
    for d, G, btype in ((2, 9, 'Pauli'), (3, 8, 'GGM'), (4, 11, 'Pauli')):
        base = rand_pulse_sequence(ff, rng, d, G, 2, 3, btype=btype)
        # grid sizes: fewer intervals than ranks, not divisible by the world size, large
        for n_omega in (1, 2, 3, world, world + 1, 64, 131, 5000):
            omega = np.geomspace(0.05, 20, n_omega) if n_omega > 1 else np.array([0.3])
            S1 = 1e-2/omega
            S2 = np.array([S1*(k + 1) for k in range(3)])
            S3 = np.einsum('a,b,o->abo', [1, 2, 3], [1, 2, 3], S1) + 0j
            small = n_omega <= 131
            for S in (S1, S2, S3):
                pulse = ff.PulseSequence.from_arrays(
                    base.c_opers, base.c_oper_identifiers, base.c_coeffs, base.n_opers,
                    base.n_oper_identifiers, base.n_coeffs, base.dt, base.basis)   # cold every time
                got = ffd.infidelity(pulse, S, omega)
                if small or S is S1:
                    check(f'infidelity d={d} n={n_omega} ndim={S.ndim}', got,
                          oracle_infidelity(base, S, omega))
                # identical bits on every rank: every rank puts its result into its own row of a zero
                # matrix; the sum over ranks (exact: one non-zero term per entry) must have equal rows
                mine = np.zeros((world,) + np.shape(got))
                mine[rank] = got
                rows = ffd.allreduce_sum(mine)
                if not all(np.array_equal(rows[r], rows[0]) for r in range(world)):
                    failures.append(('infidelity differs between ranks', d, n_omega, S.ndim))
            ids = list(base.n_oper_identifiers[[2, 0]])
            pulse = ff.PulseSequence.from_arrays(
                base.c_opers, base.c_oper_identifiers, base.c_coeffs, base.n_opers,
                base.n_oper_identifiers, base.n_coeffs, base.dt, base.basis)
            got = ffd.infidelity(pulse, S2[:2], omega, n_oper_identifiers=ids)
            if small:
                check(f'infidelity ids d={d} n={n_omega}', got, oracle_infidelity(base, S2[:2], omega, ids))
            if small:
                F = ffd.filter_function(pulse, omega)
                check(f'filter_function d={d} n={n_omega}', F, oracle_filter_function(base, omega))

        # sharded concatenation of the pieces == the whole pulse
        pieces = [base[0:3], base[3:4], base[4:G]]
        for n_omega in (1, 2, 5, 64, 131):
            omega = np.geomspace(0.05, 20, n_omega) if n_omega > 1 else np.array([0.3])
            joined, F = ffd.concatenate(pieces, omega)
            check(f'concatenate d={d} n={n_omega}', F, oracle_filter_function(base, omega), 1e-9)
            if abs(joined.tau - base.tau) > 1e-12:
                failures.append(('concatenate tau', d, n_omega))

    # the bench workloads at full size against the reference's own results (tests/golden)
    for name in ('c2', 'd4'):
        path = os.path.join(GOLDEN, f'workload_full_{name}.npz')
        if not os.path.exists(path):
            continue
        g = np.load(path)
        wl = workloads.get(name)
        pulse = ff.PulseSequence(
            [[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
            [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)], wl.dt,
            ff.Basis.pauli(int(np.log2(wl.d))))
        got = ffd.infidelity(pulse, wl.spectrum, wl.omega)
        check(f'{name} full-size infidelity vs reference', got, g['infidelity'])
        pulse.cleanup('all')
        F = ffd.filter_function(pulse, wl.omega)
        check(f'{name} full-size filter function vs reference', F[..., g['pick']],
              g['filter_function'])

    torch.cuda.synchronize()
    dist.barrier()
    flag = torch.tensor([len(failures)], device='cuda')
    dist.all_reduce(flag)
    if failures:
        print(f'rank {rank}: {len(failures)} FAILURES, first {failures[:6]}', flush=True)
    if int(flag.item()):
        dist.destroy_process_group()
        sys.exit(1)
    if rank == 0:
        print(f'DIST_GPU_OK world={world} peers={"nvlink" if peers is not None else "nccl"}',
              flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
