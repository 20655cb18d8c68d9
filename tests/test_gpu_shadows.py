"""Device-resident shadows of cached arrays (SURVEY.md 8b ``dev_handle``; the reference's cache protocol is
pulse_sequence.py:262-271, :1158-1245): consumers of a pulse's cached control matrix / filter function
must not upload them again, must give the same numbers, and the shadow must end with the cache entry."""
import gc

import numpy as np
import pytest

import ff_oracle as oracle
from helpers import nerr, rand_pulse_sequence

pytestmark = pytest.mark.gpu
TOL = 1e-10


def oracle_B(pulse, omega):
    H = oracle.hamiltonian_from_coeffs(pulse.c_opers, pulse.c_coeffs)
    ev, V, Q = oracle.diagonalize(H, pulse.dt)
    return oracle.control_matrix_from_scratch(ev, V, Q, omega, np.asarray(pulse.basis), pulse.n_opers,
                                              pulse.n_coeffs, pulse.dt)


def stats(ff):
    from filter_functions_b200 import _lib
    return _lib.shadow_stats()


def test_second_spectrum_and_decay_amplitudes_read_the_device_copy(engine):
    ff = engine
    gc.collect()
    count0, bytes0, hits0, hit_bytes0 = stats(ff)
    rng = np.random.default_rng(11)
    pulse = rand_pulse_sequence(ff, rng, 4, 30, 2, 3, btype='Pauli')
    omega = np.geomspace(1e-2, 50, 3000)
    S = 1e-3/omega**0.7
    first = ff.infidelity(pulse, S, omega)                     # cold pipeline: results stay mirrored
    count1, bytes1, hits1, _ = stats(ff)
    assert count1 == count0 + 1 and bytes1 > bytes0
    B = pulse.get_control_matrix(omega)
    F = pulse.get_filter_function(omega)
    assert not B.flags.writeable and not F.flags.writeable      # the mirror is what the library reads
    with pytest.raises(ValueError):
        B[0, 0, 0] = 0
    # second spectrum: F is not uploaded again
    S2 = np.array([S*(k + 1) for k in range(3)])
    second = ff.infidelity(pulse, S2, omega)
    _, _, hits2, hit_bytes2 = stats(ff)
    assert hits2 > hits1 and hit_bytes2 - hit_bytes0 >= F.nbytes
    B_o = oracle_B(pulse, omega)
    F_o = oracle.filter_function(B_o)
    np.testing.assert_allclose(first, oracle.infidelity_from_filter_function(F_o, S, omega, pulse.d),
                               rtol=TOL)
    np.testing.assert_allclose(second, oracle.infidelity_from_filter_function(F_o, S2, omega, pulse.d),
                               rtol=TOL)
    # decay amplitudes and the generalized filter function start from the mirrored control matrix
    _, _, hits3, hit_bytes3 = stats(ff)
    Gamma = ff.numeric.calculate_decay_amplitudes(pulse, S, omega)
    _, _, hits4, hit_bytes4 = stats(ff)
    assert hits4 > hits3 and hit_bytes4 - hit_bytes3 >= B.nbytes
    want = np.trapezoid(np.einsum('ako,o,alo->aklo', B_o.conj(), S, B_o).real, omega)/(2*np.pi)
    assert nerr(Gamma, want) < TOL
    F_gen = pulse.get_filter_function(omega, which='generalized')
    assert nerr(F_gen, np.einsum('ako,blo->abklo', B_o.conj(), B_o)) < TOL
    # a copy of the cached array is an ordinary writable array and is uploaded as usual
    B2 = B.copy()
    B2 *= 2
    assert nerr(ff.numeric.calculate_filter_function(B2), 4*F_o) < TOL
    # the shadow ends with the cache entries
    del B, F, F_gen
    pulse.cleanup('all')
    gc.collect()
    count5, bytes5, _, _ = stats(ff)
    assert count5 == count0 and bytes5 == bytes0


def test_new_frequencies_replace_the_shadow(engine):
    ff = engine
    gc.collect()
    count0 = stats(ff)[0]
    rng = np.random.default_rng(12)
    pulse = rand_pulse_sequence(ff, rng, 2, 50, 2, 2, btype='Pauli')
    om1, om2 = np.geomspace(1e-2, 50, 4000), np.geomspace(1e-1, 20, 5000)
    F1 = pulse.get_filter_function(om1).copy()
    assert stats(ff)[0] == count0 + 1
    F2 = pulse.get_filter_function(om2)          # other grid: new control matrix, new mirror; the first
    # block stays alive (and mirrored) as long as the eigensystem it also holds is cached
    assert stats(ff)[0] == count0 + 2
    assert nerr(F1, oracle.filter_function(oracle_B(pulse, om1))) < TOL
    assert nerr(F2, oracle.filter_function(oracle_B(pulse, om2))) < TOL
    np.testing.assert_allclose(ff.infidelity(pulse, 1/om2, om2),
                               oracle.infidelity_from_filter_function(F2, 1/om2, om2, pulse.d), rtol=TOL)
    del F2
    pulse.cleanup('frequency dependent')
    gc.collect()
    # eigensystem and control matrix shared one mirror with the frequency data: all gone only when the
    # time-domain arrays are dropped too
    pulse.cleanup('all')
    gc.collect()
    assert stats(ff)[0] == count0


def test_concatenation_reads_mirrored_gate_control_matrices(engine):
    """Config-5 situation: gates with cached control matrices and differing noise operators are
    concatenated; the cached rows travel device-to-device."""
    ff = engine
    gc.collect()
    rng = np.random.default_rng(13)
    d = 4
    omega = np.geomspace(1e-2, 30, 2500)
    basis = ff.Basis.pauli(2)

    def herm(n):
        A = rng.standard_normal((n, d, d)) + 1j*rng.standard_normal((n, d, d))
        return (A + A.conj().transpose(0, 2, 1))/2

    ops = herm(4)
    noise = herm(3)
    gates = []
    for i in range(3):
        G = int(rng.integers(2, 5))
        H_c = [[ops[k], rng.standard_normal(G), f'c{k}'] for k in range(2)]
        ids = [0, 1] if i == 0 else [1, 2] if i == 1 else [0, 2]      # every gate lacks one operator
        H_n = [[noise[k], np.ones(G), f'n{k}'] for k in ids]
        gate = ff.PulseSequence(H_c, H_n, 1 - rng.random(G)*0.5, basis)
        gate.cache_control_matrix(omega)
        gates.append(gate)
    hits0, hit_bytes0 = stats(ff)[2:]
    total = ff.concatenate(gates, omega=omega)
    hits1, hit_bytes1 = stats(ff)[2:]
    cached_bytes = sum(g.get_control_matrix(omega).nbytes for g in gates)
    assert hits1 > hits0 and hit_bytes1 - hit_bytes0 >= cached_bytes
    whole = ff.concatenate(gates, calc_filter_function=False)
    assert nerr(total.get_control_matrix(omega), oracle_B(whole, omega)) < 1e-9
    assert nerr(total.get_filter_function(omega),
                oracle.filter_function(oracle_B(whole, omega))) < 1e-9
    # the concatenated pulse's own arrays are mirrored as well: its infidelity starts on the device
    hits2 = stats(ff)[2]
    ff.infidelity(total, 1e-2/omega, omega)
    assert stats(ff)[2] > hits2
