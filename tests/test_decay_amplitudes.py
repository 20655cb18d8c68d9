"""Decay amplitudes, cumulant function and error transfer matrix.

Reference: numeric.calculate_decay_amplitudes (numeric.py:1194-1337), calculate_cumulant_function
(:957-1191), error_transfer_matrix (:1938-2059); reference tests tests/test_precision.py:631-727 and
tests/test_core.py:808-994.  The fixture tests/golden/decay_amplitudes.npz was generated from the
reference itself by oracle/gen_golden.py.  CPU tests pin the oracle and the host-side trace
contraction; the ``gpu`` tests compare the CUDA path with the oracle and the fixture (rtol 1e-10 as
normalised max-abs error, atol 1e-14 as in the reference's own tests)."""
import os

import numpy as np
import pytest

import ff_oracle as oracle
from helpers import nerr, rand_pulse_sequence
from test_oracle import GOLDEN

TOL = 1e-10
CASES = [(2, 'Pauli'), (3, 'GGM'), (4, 'Pauli')]


def fixture():
    return np.load(os.path.join(GOLDEN, 'decay_amplitudes.npz'))


def oracle_control_matrix(g, tag, sl=slice(None)):
    H = oracle.hamiltonian_from_coeffs(g[f'{tag}_c_opers'], g[f'{tag}_c_coeffs'][:, sl])
    dt = g[f'{tag}_dt'][sl]
    ev, V, Q = oracle.diagonalize(H, dt)
    B = oracle.control_matrix_from_scratch(ev, V, Q, g[f'{tag}_omega'], g[f'{tag}_basis'],
                                           g[f'{tag}_n_opers'], g[f'{tag}_n_coeffs'][:, sl], dt)
    return B, Q[-1]


@pytest.mark.parametrize('d,btype', CASES)
def test_oracle_against_reference_fixture(d, btype):
    g, tag = fixture(), f'd{d}'
    B, _ = oracle_control_matrix(g, tag)
    omega = g[f'{tag}_omega']
    for i in range(3):
        S = g[f'{tag}_spectrum{i}']
        Gamma = oracle.decay_amplitudes(B, S, omega)
        assert nerr(Gamma, g[f'{tag}_decay_amplitudes{i}']) < 1e-12
        K = oracle.cumulant_function(Gamma, g[f'{tag}_basis'], btype)
        assert nerr(K, g[f'{tag}_cumulant_function{i}']) < 1e-12
        assert nerr(oracle.error_transfer_matrix(K), g[f'{tag}_error_transfer_matrix{i}']) < 1e-12
    # pulse correlations: split pulse, concatenate by hand (B = B1 + phase B2 Q1)
    split, G = int(g[f'{tag}_split']), len(g[f'{tag}_dt'])
    B1, U1 = oracle_control_matrix(g, tag, slice(0, split))
    B2, _ = oracle_control_matrix(g, tag, slice(split, G))
    phases = oracle.total_phases(omega, g[f'{tag}_dt'][:split].sum())[None]
    L1 = oracle.liouville_representation(U1[None], g[f'{tag}_basis']).real
    B_pc = oracle.control_matrix_from_atomic(phases, np.stack([B1, B2]), L1, 'correlations')
    assert nerr(B_pc.sum(0), B) < 1e-12
    for i in range(3):
        Gamma_pc = oracle.decay_amplitudes(B_pc, g[f'{tag}_spectrum{i}'], omega)
        assert nerr(Gamma_pc, g[f'{tag}_decay_amplitudes_pc{i}']) < 1e-12
        assert nerr(Gamma_pc.sum(axis=(0, 1)), g[f'{tag}_decay_amplitudes{i}']) < 1e-12


@pytest.mark.parametrize('d,btype', CASES + [(5, 'GGM')])
def test_trace_tensor_free_contraction(d, btype):
    """The product's O(n_basis^2 d^2) contraction equals the reference's n_basis^4 trace-tensor one."""
    import filter_functions_b200 as ff
    from filter_functions_b200 import numeric
    rng = np.random.default_rng(d)
    basis = ff.Basis.pauli(int(np.log2(d))) if btype == 'Pauli' else ff.Basis.ggm(d)
    Gamma = rng.standard_normal((2, 3, len(basis), len(basis)))

    class FakePulse:
        pass
    pulse = FakePulse()
    pulse.basis = basis
    K = numeric.calculate_cumulant_function(pulse, decay_amplitudes=Gamma)
    assert K.shape == Gamma.shape
    assert nerr(K, oracle.cumulant_function(Gamma, np.asarray(basis), btype)) < 1e-13
    etm = ff.error_transfer_matrix(cumulant_function=K)
    assert nerr(etm, oracle.error_transfer_matrix(K)) < 1e-13
    with pytest.raises(ValueError):
        numeric.calculate_cumulant_function(pulse)
    with pytest.raises(ValueError):
        ff.error_transfer_matrix(pulse)
    with pytest.raises(NotImplementedError):
        numeric.calculate_cumulant_function(pulse, decay_amplitudes=Gamma, second_order=True)
    with pytest.raises(TypeError):
        ff.error_transfer_matrix(cumulant_function=[1, 2, 3])
    with pytest.raises(ValueError):
        ff.error_transfer_matrix(cumulant_function=K[0, 0, :, :3])


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
def pulse_from_fixture(ff, g, tag, btype, sl=slice(None)):
    basis = ff.Basis(g[f'{tag}_basis'], btype=btype)
    return ff.PulseSequence(
        [[op, c[sl], str(i)] for op, c, i in zip(g[f'{tag}_c_opers'], g[f'{tag}_c_coeffs'],
                                                 g[f'{tag}_c_ids'])],
        [[op, c[sl], str(i)] for op, c, i in zip(g[f'{tag}_n_opers'], g[f'{tag}_n_coeffs'],
                                                 g[f'{tag}_n_ids'])],
        g[f'{tag}_dt'][sl], basis)


@pytest.mark.gpu
@pytest.mark.parametrize('d,btype', CASES)
def test_fixture_through_public_api(engine, d, btype):
    ff = engine
    g, tag = fixture(), f'd{d}'
    omega = g[f'{tag}_omega']
    pulse = pulse_from_fixture(ff, g, tag, btype)
    for i in range(3):
        S = g[f'{tag}_spectrum{i}']
        Gamma = ff.numeric.calculate_decay_amplitudes(pulse, S, omega)
        ref = g[f'{tag}_decay_amplitudes{i}']
        assert Gamma.shape == ref.shape and Gamma.dtype == np.float64
        np.testing.assert_allclose(Gamma, ref, rtol=TOL, atol=TOL*np.abs(ref).max())
        K = ff.numeric.calculate_cumulant_function(pulse, S, omega)
        ref = g[f'{tag}_cumulant_function{i}']
        np.testing.assert_allclose(K, ref, rtol=TOL, atol=TOL*np.abs(ref).max())
        etm = ff.error_transfer_matrix(pulse, S, omega)
        np.testing.assert_allclose(etm, g[f'{tag}_error_transfer_matrix{i}'], rtol=TOL, atol=1e-14)
    ids = [str(i) for i in g[f'{tag}_subset_ids']]
    n_nops = len(g[f'{tag}_n_ids'])
    sub = ff.numeric.calculate_decay_amplitudes(pulse, g[f'{tag}_spectrum1'][[0, n_nops - 1]],
                                                omega, n_oper_identifiers=ids)
    ref = g[f'{tag}_decay_amplitudes_subset']
    np.testing.assert_allclose(sub, ref, rtol=TOL, atol=TOL*np.abs(ref).max())
    # generalized filter function cached, control matrix dropped: same integral from F_gen
    F_gen = pulse.get_filter_function(omega, which='generalized')
    fresh = pulse_from_fixture(ff, g, tag, btype)
    fresh.cache_filter_function(omega, filter_function=F_gen, which='generalized')
    assert not fresh.is_cached('control_matrix')
    for i in range(3):
        Gamma = ff.numeric.calculate_decay_amplitudes(fresh, g[f'{tag}_spectrum{i}'], omega)
        ref = g[f'{tag}_decay_amplitudes{i}']
        np.testing.assert_allclose(Gamma, ref, rtol=TOL, atol=TOL*np.abs(ref).max())
    # pulse correlations
    split, G = int(g[f'{tag}_split']), len(g[f'{tag}_dt'])
    halves = [pulse_from_fixture(ff, g, tag, btype, slice(0, split)),
              pulse_from_fixture(ff, g, tag, btype, slice(split, G))]
    for h in halves:
        h.cache_control_matrix(omega)
    seq = ff.concatenate(halves, calc_pulse_correlation_FF=True)
    for i in range(3):
        Gamma_pc = ff.numeric.calculate_decay_amplitudes(seq, g[f'{tag}_spectrum{i}'], omega,
                                                         which='correlations')
        ref = g[f'{tag}_decay_amplitudes_pc{i}']
        assert Gamma_pc.shape == ref.shape
        np.testing.assert_allclose(Gamma_pc, ref, rtol=TOL, atol=TOL*np.abs(ref).max())
    K_pc = ff.numeric.calculate_cumulant_function(seq, g[f'{tag}_spectrum0'], omega,
                                                  which='correlations')
    ref = g[f'{tag}_cumulant_function_pc0']
    np.testing.assert_allclose(K_pc, ref, rtol=TOL, atol=TOL*np.abs(ref).max())
    with pytest.raises(ValueError):
        ff.numeric.calculate_decay_amplitudes(seq, g[f'{tag}_spectrum0'][:-1], omega[:-1],
                                              which='correlations')
    with pytest.raises(ff.util.CalculationError):
        ff.numeric.calculate_decay_amplitudes(pulse, g[f'{tag}_spectrum0'], omega,
                                              which='correlations')


@pytest.mark.gpu
@pytest.mark.parametrize('n_nops,n_basis,n_omega,P', [
    (1, 4, 1, 1), (2, 4, 2, 1), (3, 4, 7, 1), (3, 9, 130, 1), (2, 15, 257, 1), (6, 16, 3001, 1),
    (2, 36, 515, 1), (1, 81, 300, 1), (1, 256, 700, 1), (2, 4, 301, 3), (2, 16, 97, 2)])
def test_kernel_against_oracle_random_control_matrices(engine, n_nops, n_basis, n_omega, P):
    """The kernel on arbitrary complex arrays (ragged sizes: n_basis not a multiple of 8, n_omega not
    a multiple of the frequency block, a single frequency -> zero integral), all spectrum shapes."""
    ff = engine
    from filter_functions_b200 import numeric
    rng = np.random.default_rng(n_basis*1000 + n_omega)
    shape = (n_nops, n_basis, n_omega) if P == 1 else (P, n_nops, n_basis, n_omega)
    B = rng.standard_normal(shape) + 1j*rng.standard_normal(shape)
    omega = np.sort(rng.random(n_omega))*10 + 0.1
    S1 = 1/omega
    S2 = np.array([S1*(i + 1) for i in range(n_nops)])
    A = rng.standard_normal((n_nops, n_nops, n_omega)) + 1j*rng.standard_normal((n_nops, n_nops,
                                                                                 n_omega))
    S3 = A + A.conj().swapaxes(0, 1)
    idx = np.arange(n_nops)
    for S in (S1, S2, S3):
        got = numeric._decay_amplitudes_from_control_matrix(B, S, omega, idx)
        ref = oracle.decay_amplitudes(B, S, omega)
        assert got.shape == ref.shape
        assert nerr(got, ref) < 1e-12 or np.abs(ref).max() == 0
        if n_omega == 1:
            assert not got.any()
    if n_nops > 1:
        sel = np.array([n_nops - 1, 0])
        got = numeric._decay_amplitudes_from_control_matrix(B, S3[sel][:, sel], omega, sel)
        ref = oracle.decay_amplitudes(B, S3[sel][:, sel], omega, sel)
        assert nerr(got, ref) < 1e-12


@pytest.mark.gpu
def test_consistency_with_infidelity_and_generalized_filter_function(engine):
    """tr(Gamma)/d is the infidelity (numeric.py:2277-2285 of the reference docs), and Gamma is the
    integral of the generalized filter function; d = 8 also exercises the trace-free contraction
    behind a real pulse."""
    ff = engine
    rng = np.random.default_rng(99)
    for d, btype in ((2, 'Pauli'), (3, 'GGM'), (8, 'Pauli')):
        pulse = rand_pulse_sequence(ff, rng, d, 12, 2, 3, btype=btype)
        omega = np.geomspace(0.01, 50, 400)
        S = 1e-3/omega
        Gamma = ff.numeric.calculate_decay_amplitudes(pulse, S, omega)
        infid = ff.infidelity(pulse, S, omega)
        np.testing.assert_allclose(np.einsum('akk->a', Gamma)/d, infid, rtol=1e-10)
        F_gen = pulse.get_filter_function(omega, which='generalized')
        ref = oracle.integrate((F_gen[range(3), range(3)]*S).real, omega)/(2*np.pi)
        assert nerr(Gamma, ref) < 1e-11
        K = ff.numeric.calculate_cumulant_function(pulse, S, omega)
        if d <= 3:
            assert nerr(K, oracle.cumulant_function(Gamma, np.asarray(pulse.basis), btype)) < 1e-11
        # the identity row and column vanish for a traceless basis; -tr(K)/d^2 is the infidelity
        # exactly, 1 - tr(exp K)/d^2 to first order in the noise strength
        assert np.abs(K[..., 0, :]).max() < 1e-13*max(1, np.abs(K).max())
        np.testing.assert_allclose(-np.trace(K.sum(axis=0))/d**2, infid.sum(), rtol=1e-9)
        etm = ff.error_transfer_matrix(pulse, S, omega)
        np.testing.assert_allclose(1 - np.trace(etm)/d**2, infid.sum(), rtol=5e-2)
