"""Periodic repetition: numeric.calculate_control_matrix_periodic
(numeric.py:884-954) and concatenate_periodic (pulse_sequence.py:1890-1977); modelled on the
reference's tests/test_sequencing.py:608-688."""
from itertools import repeat

import numpy as np
import pytest

import ff_oracle as oracle
from helpers import nerr, rand_pulse_sequence

TOL = 1e-10


def test_oracle_periodic_equals_explicit_sum():
    rng = np.random.default_rng(8)
    n_nops, n, n_omega = 2, 9, 21
    B = rng.standard_normal((n_nops, n, n_omega)) + 1j*rng.standard_normal((n_nops, n, n_omega))
    Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
    phases = oracle.cexp(rng.random(n_omega)*7)
    phases[3] = 1.0                         # I - T singular if Q has eigenvalue 1
    Qm[:, 0], Qm[0, :] = 0, 0
    Qm[0, 0] = 1.0
    for G in (1, 2, 5, 12):
        atomic = np.broadcast_to(B, (G,) + B.shape)
        ph = np.cumprod(np.broadcast_to(phases, (G - 1, n_omega)), axis=0)
        Qs = np.array([np.linalg.matrix_power(Qm, g + 1) for g in range(G - 1)])
        ref = oracle.control_matrix_from_atomic(ph, atomic, Qs) if G > 1 else B
        assert nerr(oracle.control_matrix_periodic(phases, B, Qm, G), ref) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize('n_nops,n_basis,n_omega', [(1, 4, 33), (3, 9, 300), (2, 16, 1001),
                                                    (1, 36, 257), (2, 256, 40)])
def test_kernel_against_oracle(engine, n_nops, n_basis, n_omega):
    """Random control matrices, orthogonal L with an eigenvalue exactly 1 (where the reference has to
    leave its linear solve), phases hitting 1 exactly; repeats = 1, primes, powers of two, 10^4."""
    rng = np.random.default_rng(n_basis + n_omega)
    shape = (n_nops, n_basis, n_omega)
    B = rng.standard_normal(shape) + 1j*rng.standard_normal(shape)
    Qm, _ = np.linalg.qr(rng.standard_normal((n_basis, n_basis)))
    Qm[:, 0], Qm[0, :] = 0, 0
    Qm[0, 0] = 1.0
    phases = oracle.cexp(rng.random(n_omega)*50)
    phases[0] = 1.0
    f = engine.numeric.calculate_control_matrix_periodic
    for G in (1, 2, 3, 16, 97, 1000) + ((10_000,) if n_basis <= 16 else ()):
        got = f(phases, B, Qm, G)
        ref = oracle.control_matrix_periodic(phases, B, Qm, G)
        assert got.shape == ref.shape
        assert nerr(got, ref) < TOL, G
    # complex Liouville propagator (non-Hermitian basis)
    Qc = Qm*np.exp(0.3j)
    assert nerr(f(phases, B, Qc, 13), oracle.control_matrix_periodic(phases, B, Qc, 13)) < TOL
    with pytest.raises(ValueError):
        f(phases, B, Qm, 0)
    with pytest.raises(ValueError):
        f(phases, B[0], Qm, 2)


@pytest.mark.gpu
def test_concatenate_periodic_rotating_frame_not_gate(engine):
    """The reference's driven-qubit example (tests/test_sequencing.py:608-668): one period cached,
    repeated G times == the whole pulse from scratch == ff.concatenate of G copies."""
    ff = engine
    X, Y, Z = ff.util.paulis[1:]
    A, omega_0 = 0.01, 1.0
    tau = np.pi/A
    omega = np.logspace(-3, 3, 501)
    t = np.linspace(0, tau, 1001)
    dt = np.diff(t)
    lab = ff.PulseSequence([[Z, [omega_0/2]*len(dt)], [X, A*np.cos(omega_0*t[1:])]],
                           [[Z, np.ones_like(dt)], [X, np.ones_like(dt)]], dt)
    F_lab = lab.get_filter_function(omega)
    T = 2*np.pi/omega_0
    G = round(tau/T)
    t = np.linspace(0, T, int(T/lab.dt[0]) + 1)
    dt = np.diff(t)
    atomic = ff.PulseSequence([[Z, [omega_0/2]*len(dt)], [X, A*np.cos(omega_0*t[1:])]],
                              [[Z, np.ones_like(dt)], [X, np.ones_like(dt)]], dt)
    atomic.cache_filter_function(omega)
    cc = ff.concatenate(atomic for _ in range(G))
    periodic = ff.concatenate_periodic(atomic, G)
    for attr in ('dt', 'c_opers', 'c_coeffs', 'n_opers', 'n_coeffs'):
        np.testing.assert_allclose(getattr(lab, attr), getattr(periodic, attr), atol=1e-15, rtol=1e2)
    for pulse in (cc, periodic):
        assert 'total_phases' in pulse.frequency_data
        assert 'total_propagator' in pulse.data and 'total_propagator_liouville' in pulse.data
    np.testing.assert_allclose(F_lab, cc.get_filter_function(omega), atol=1e-13, rtol=1e-7)
    np.testing.assert_allclose(F_lab, periodic.get_filter_function(omega), atol=1e-13, rtol=1e-7)
    assert nerr(periodic.get_control_matrix(omega), cc.get_control_matrix(omega)) < TOL
    assert nerr(periodic.total_propagator, cc.total_propagator) < 1e-12
    with pytest.raises(TypeError):
        ff.concatenate_periodic([atomic], 2)
    bare = ff.concatenate_periodic(ff.PulseSequence([[X, [1.0]]], [[Z, [1.0]]], [1.0]), 3)
    assert len(bare) == 3 and not bare.is_cached('control_matrix') and bare.tau == 3.0


@pytest.mark.gpu
def test_concatenate_periodic_random(engine):
    """Reference tests/test_sequencing.py:670-688: random pulses, random repeat counts incl. 1."""
    ff = engine
    rng = np.random.default_rng(2718)
    for d, G in zip(rng.integers(2, 7, 8), np.concatenate([[1], rng.integers(2, 300, 7)])):
        pulse = rand_pulse_sequence(ff, rng, int(d), 5, 2, 2)
        pulse.cache_filter_function(rng.random(37))
        a = ff.concatenate(repeat(pulse, int(G)))
        b = ff.concatenate_periodic(pulse, int(G))
        assert a == b
        assert nerr(b.frequency_data['control_matrix'], a.frequency_data['control_matrix']) < TOL
        assert nerr(b.frequency_data['filter_function'], a.frequency_data['filter_function']) < TOL
        cm = ff.numeric.calculate_control_matrix_periodic(
            pulse.get_total_phases(pulse.omega), pulse.get_control_matrix(pulse.omega),
            pulse.total_propagator_liouville, int(G), check_invertible=False)
        assert nerr(cm, a.frequency_data['control_matrix']) < TOL
