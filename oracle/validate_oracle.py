"""Pin ``oracle/ff_oracle.py`` against the UNMODIFIED reference imported from /root/reference.

Runs only in the build container (the reference is not shipped to the GPU box).  Usage::

    python oracle/validate_oracle.py

Exits non-zero if any gauge-invariant quantity of the oracle deviates from the reference by more
than 1e-12 (normalised max-abs).  The reference needs ``opt_einsum`` and ``sparse`` which are not
installed here; ``oracle/shim`` provides NumPy-backed stand-ins.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, '/root/reference/tests')
sys.path.insert(0, HERE)

import filter_functions as ff  # noqa: E402  (the reference)
from filter_functions import numeric as ref_numeric  # noqa: E402
from filter_functions import superoperator as ref_super  # noqa: E402
from filter_functions import util as ref_util  # noqa: E402

import ff_oracle as oracle  # noqa: E402


def nerr(x, ref):
    ref = np.asarray(ref)
    scale = np.abs(ref).max()
    return np.abs(np.asarray(x) - ref).max()/(scale if scale > 0 else 1.0)


def rand_herm_traceless(rng, d, n):
    A = rng.standard_normal((n, d, d)) + 1j*rng.standard_normal((n, d, d))
    A = (A + A.conj().transpose(0, 2, 1))/2
    A -= np.einsum('njj->n', A)[:, None, None]*np.eye(d)/d
    return A


def main():
    rng = np.random.default_rng(20261017)
    worst = {}

    def record(name, err):
        worst[name] = max(worst.get(name, 0.0), float(err))

    for d, G, n_nops, btype in [(2, 30, 3, 'pauli'), (3, 17, 2, 'ggm'), (4, 25, 6, 'pauli'),
                                (5, 9, 2, 'ggm'), (8, 6, 3, 'pauli')]:
        c_opers = rand_herm_traceless(rng, d, 3)
        n_opers = rand_herm_traceless(rng, d, n_nops)
        c_coeffs = rng.standard_normal((3, G))
        n_coeffs = rng.random((n_nops, G)) + 0.5
        dt = 1 - rng.random(G)
        basis = ff.Basis.pauli(int(np.log2(d))) if btype == 'pauli' else ff.Basis.ggm(d)
        my_basis = oracle.pauli_basis(int(np.log2(d))) if btype == 'pauli' else oracle.ggm_basis(d)
        record('basis', nerr(my_basis, np.asarray(basis)))

        omega = np.concatenate(([0.0], np.geomspace(1e-3, 50, 60), -np.geomspace(1e-2, 5, 7)))
        H = np.einsum('ijk,il->ljk', c_opers, c_coeffs)
        record('hamiltonian', nerr(oracle.hamiltonian_from_coeffs(c_opers, c_coeffs), H))

        ev_r, V_r, Q_r = ref_numeric.diagonalize(H, dt)
        ev_o, V_o, Q_o = oracle.diagonalize(H, dt)
        record('eigvals', nerr(ev_o, ev_r))
        record('propagators', nerr(Q_o, Q_r))

        t = np.concatenate(([0], dt.cumsum()))
        B_r = ref_numeric.calculate_control_matrix_from_scratch(
            ev_r, V_r, Q_r, omega, basis, n_opers, n_coeffs, dt, t)
        B_o = oracle.control_matrix_from_scratch(ev_r, V_r, Q_r, omega, my_basis, n_opers,
                                                 n_coeffs, dt, t)
        record('control_matrix', nerr(B_o, B_r))

        if d <= 5:
            B_ri, inter_r = ref_numeric.calculate_control_matrix_from_scratch(
                ev_r, V_r, Q_r, omega, basis, n_opers, n_coeffs, dt, t, cache_intermediates=True)
            B_oi, inter_o = oracle.control_matrix_intermediates(ev_r, V_r, Q_r, omega, my_basis,
                                                                n_opers, n_coeffs, dt, t)
            record('intermediates_control_matrix', nerr(B_oi, B_ri))
            assert sorted(inter_r) == sorted(inter_o)
            for key in inter_r:
                record('intermediates_' + key, nerr(inter_o[key], inter_r[key]))

        for which in ('fidelity', 'generalized'):
            record('filter_function_' + which,
                   nerr(oracle.filter_function(B_r, which),
                        ref_numeric.calculate_filter_function(B_r, which)))

        L_r = ref_super.liouville_representation(Q_r[1:4], basis)
        record('liouville', nerr(oracle.liouville_representation(Q_r[1:4], my_basis), L_r))

        # un-normalised (rescaled) basis: the expansion divides by tr(C_j C_j)
        scaled = ff.Basis(np.asarray(basis)*np.linspace(0.5, 2.0, len(basis))[:, None, None])
        record('liouville_unnormalised',
               nerr(oracle.liouville_representation(Q_r[1:4], np.asarray(scaled)),
                    ref_super.liouville_representation(Q_r[1:4], scaled)))

        # concatenation of three "pulses" whose control matrices are random
        P = 4
        atomic = (rng.standard_normal((P, n_nops, len(my_basis), len(omega)))
                  + 1j*rng.standard_normal((P, n_nops, len(my_basis), len(omega))))
        phases = ref_util.cexp(np.outer(rng.random(P - 1)*3, omega))
        Lq = ref_super.liouville_representation(Q_r[2:2 + P - 1], basis)
        for which in ('total', 'correlations'):
            record('from_atomic_' + which,
                   nerr(oracle.control_matrix_from_atomic(phases, atomic, Lq, which),
                        ref_numeric.calculate_control_matrix_from_atomic(phases, atomic, Lq,
                                                                         which=which)))
        for repeats in (1, 2, 7, 64):
            record('control_matrix_periodic',
                   nerr(oracle.control_matrix_periodic(phases[0], atomic[0], Lq[0], repeats),
                        ref_numeric.calculate_control_matrix_periodic(phases[0], atomic[0], Lq[0],
                                                                      repeats)))
        for which in ('fidelity', 'generalized'):
            record('pc_filter_function_' + which,
                   nerr(oracle.pulse_correlation_filter_function(atomic, which),
                        ref_numeric.calculate_pulse_correlation_filter_function(atomic, which)))

        # infidelity through the reference's public API vs oracle on the reference's F
        pulse = ff.PulseSequence(list(zip(c_opers, c_coeffs)), list(zip(n_opers, n_coeffs)), dt,
                                 basis)
        om = np.geomspace(0.1, 10, 51)
        F = pulse.get_filter_function(om)
        S1 = 1e-2/om
        S2 = np.array([S1*(i + 1) for i in range(n_nops)])
        S3 = np.einsum('a,b,o->abo', np.arange(1, n_nops + 1), np.arange(1, n_nops + 1), S1)
        S3 = S3 + 1j*np.triu(np.ones((n_nops, n_nops)), 1)[..., None]*om \
            - 1j*np.tril(np.ones((n_nops, n_nops)), -1)[..., None]*om
        for S in (S1, S2, S3):
            record(f'infidelity_{S.ndim}d',
                   nerr(oracle.infidelity_from_filter_function(F, S, om, d),
                        ff.infidelity(pulse, S, om)))
            if d <= 5:   # dense n_basis^4 trace tensor in the oracle and the stand-in
                for B_in, which in ((pulse.get_control_matrix(om), 'total'),):
                    Gam_r = ref_numeric.calculate_decay_amplitudes(pulse, S, om, which=which)
                    Gam_o = oracle.decay_amplitudes(B_in, S, om)
                    record(f'decay_amplitudes_{S.ndim}d', nerr(Gam_o, Gam_r))
                    K_r = ref_numeric.calculate_cumulant_function(pulse, S, om)
                    K_o = oracle.cumulant_function(Gam_r, my_basis,
                                                   'Pauli' if btype == 'pauli' else 'GGM')
                    record(f'cumulant_function_{S.ndim}d', nerr(K_o, K_r))
                    record(f'error_transfer_matrix_{S.ndim}d',
                           nerr(oracle.error_transfer_matrix(K_r),
                                ff.error_transfer_matrix(pulse, S, om)))
        if d <= 3:
            # pulse-correlation decay amplitudes of a two-pulse sequence
            halves = [ff.PulseSequence(list(zip(c_opers, c_coeffs[:, sl])),
                                       list(zip(n_opers, n_coeffs[:, sl])), dt[sl], basis)
                      for sl in (slice(0, G//2), slice(G//2, G))]
            for h in halves:
                h.cache_control_matrix(om)
            seq = ff.concatenate(halves, calc_pulse_correlation_FF=True)
            for S in (S1, S2, S3):
                Gam_r = ref_numeric.calculate_decay_amplitudes(seq, S, om, which='correlations')
                Gam_o = oracle.decay_amplitudes(seq.get_pulse_correlation_control_matrix(), S, om)
                record(f'decay_amplitudes_pc_{S.ndim}d', nerr(Gam_o, Gam_r))
        record('sample_frequencies',
               nerr(oracle.sample_frequencies(pulse.tau, pulse.dt.min()),
                    ref_util.get_sample_frequencies(pulse)))

    x = rng.standard_normal(1000)*100
    record('cexp', nerr(oracle.cexp(x), ref_util.cexp(x)))
    record('cexpm1', nerr(oracle.cexpm1(x), ref_util.cexpm1(x)))
    f = rng.standard_normal((3, 200))
    xs = np.sort(rng.random(200))
    record('integrate', nerr(oracle.integrate(f, xs), ref_util.integrate(f, xs)))

    ok = True
    for name, err in sorted(worst.items()):
        flag = 'ok' if err < 1e-12 else 'FAIL'
        ok &= err < 1e-12
        print(f'{name:32s} {err:9.2e}  {flag}')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
