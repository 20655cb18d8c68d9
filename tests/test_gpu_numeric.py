"""GPU parity of the numeric layer (C ABI called through filter_functions_b200.numeric) against the
CPU oracle on the same seeded inputs.  Tolerance: north_star's rtol 1e-10 on gauge-invariant
quantities, as normalised max-abs error."""
import numpy as np
import pytest

import ff_oracle as oracle
from helpers import nerr, rand_herm, rand_herm_traceless

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _setup(rng, d, G, n_nops, n_cops=3):
    c_opers = rand_herm_traceless(rng, d, n_cops)
    n_opers = rand_herm_traceless(rng, d, n_nops)
    c_coeffs = rng.standard_normal((n_cops, G))
    n_coeffs = rng.random((n_nops, G)) + 0.25
    dt = 1 - rng.random(G)
    H = oracle.hamiltonian_from_coeffs(c_opers, c_coeffs)
    return c_opers, c_coeffs, n_opers, n_coeffs, dt, H


@pytest.mark.parametrize('d,G', [(2, 1), (2, 33), (3, 10), (4, 64), (5, 7), (8, 19), (16, 5),
                                 (2, 1000), (17, 3), (24, 4), (32, 3)])
def test_diagonalize(engine, d, G):
    rng = np.random.default_rng(100 + d*1000 + G)
    *_, dt, H = _setup(rng, d, G, 2)
    ev, V, Q = engine.numeric.diagonalize(H, dt)
    ev_o, V_o, Q_o = oracle.diagonalize(H, dt)
    assert ev.shape == (G, d) and V.shape == (G, d, d) and Q.shape == (G + 1, d, d)
    assert nerr(ev, ev_o) < TOL
    assert nerr(Q, Q_o) < TOL
    # eigenvectors are gauge dependent: check the characteristic equation and unitarity instead
    recon = V.conj().transpose(0, 2, 1) @ H @ V
    diag = np.zeros_like(recon)
    diag[:, range(d), range(d)] = ev
    assert nerr(recon, diag) < 1e-12
    assert nerr(V.conj().transpose(0, 2, 1) @ V, np.broadcast_to(np.eye(d), (G, d, d))) < 1e-12
    assert (np.diff(ev, axis=1) >= 0).all()


@pytest.mark.parametrize('d,G', [(2, 30011), (3, 13007), (4, 12400)])
def test_diagonalize_many_segments(engine, d, G):
    """More scan blocks than the apply kernel combines on its own (> 96 blocks of 256 / 128 segments):
    the block totals go through a second scan level.  Small rotation angles keep the long product
    well conditioned for the comparison."""
    rng = np.random.default_rng(d*G)
    *_, dt, H = _setup(rng, d, G, 2)
    dt = dt*0.05
    ev, V, Q = engine.numeric.diagonalize(H, dt)
    ev_o, V_o, Q_o = oracle.diagonalize(H, dt)
    assert nerr(ev, ev_o) < TOL
    assert nerr(Q, Q_o) < TOL


def test_diagonalize_degenerate_and_zero(engine):
    """H_c = 0 segments and exactly degenerate spectra (gauge freedom in whole subspaces)."""
    d, G = 4, 6
    H = np.zeros((G, d, d), dtype=complex)
    H[1] = np.diag([1.0, 1.0, -1.0, -1.0])
    H[2] = np.kron(np.array([[0, 1], [1, 0]]), np.eye(2))
    H[4] = np.diag([3.0, 1.0, 2.0, 1.0])
    dt = np.array([0.3, 1.0, 0.5, 0.0, 2.0, 1e-9])
    ev, V, Q = engine.numeric.diagonalize(H, dt)
    ev_o, _, Q_o = oracle.diagonalize(H, dt)
    assert nerr(ev, ev_o) < TOL and nerr(Q, Q_o) < TOL


@pytest.mark.parametrize('d', [2, 4, 6])
def test_diagonalize_reports_non_convergence(engine, d):
    """numpy.linalg.eigh raises LinAlgError when the iteration does not converge (NaN input); the
    library counts the matrices whose Jacobi iteration hit the sweep limit and the synchronous entry
    point raises util.CalculationError.  The next call is unaffected (the counter is reset)."""
    ff = engine
    rng = np.random.default_rng(3)
    H = rand_herm(rng, d, 9)
    dt = np.full(9, 0.5)
    bad = H.copy()
    bad[4, 1, 0] = np.nan
    with pytest.raises(ff.util.CalculationError):
        ff.numeric.diagonalize(bad, dt)
    ev, V, Q = ff.numeric.diagonalize(H, dt)
    assert nerr(ev, np.linalg.eigvalsh(H)) < TOL


@pytest.mark.parametrize('d,G,n_nops,btype,n_omega', [
    (2, 1, 1, 'pauli', 1), (2, 2, 1, 'pauli', 300), (2, 37, 3, 'pauli', 257), (2, 50, 2, 'ggm', 64),
    (3, 21, 2, 'ggm', 100), (4, 40, 6, 'pauli', 203), (4, 9, 3, 'ggm', 77), (5, 6, 2, 'ggm', 50),
    (6, 5, 2, 'ggm', 40), (8, 7, 2, 'pauli', 33), (16, 3, 2, 'ggm', 24), (2, 1003, 3, 'pauli', 129),
    (17, 2, 1, 'ggm', 20), (24, 2, 1, 'ggm', 16), (32, 2, 1, 'pauli', 12),   # up to the documented d <= 32
])
def test_control_matrix_from_scratch(engine, d, G, n_nops, btype, n_omega):
    rng = np.random.default_rng(7 + 31*d + G)
    _, _, n_opers, n_coeffs, dt, H = _setup(rng, d, G, n_nops)
    ev, V, Q = oracle.diagonalize(H, dt)
    basis = oracle.pauli_basis(int(np.log2(d))) if btype == 'pauli' else oracle.ggm_basis(d)
    omega = np.geomspace(1e-3, 60, n_omega) if n_omega > 1 else np.array([0.7])
    t = np.concatenate(([0], dt.cumsum()))
    B = engine.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers,
                                                             n_coeffs, dt, t)
    B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs, dt, t)
    assert B.shape == (n_nops, len(basis), n_omega) and B.dtype == np.complex128
    assert B.flags.c_contiguous
    for j in range(n_nops):  # per noise operator, as BASELINE.md section 3 prescribes
        assert nerr(B[j], B_o[j]) < TOL


@pytest.mark.parametrize('d,n_nops,btype', [(2, 2, 'pauli'), (2, 3, 'pauli'), (2, 4, 'pauli'),
                                            (3, 1, 'ggm'), (2, 3, 'shuffled'), (2, 3, 'scaled')])
def test_control_matrix_identity_basis_element(engine, d, n_nops, btype):
    """Thread-per-frequency kernel with the rows of an identity basis element split off (they carry
    no level-pair coefficients).  The noise operators here are NOT traceless, so those rows are
    non-zero: B_j0 = tr(B_j)/sqrt(d) sum_g s_j e^{i w t_g} I(w).  'shuffled' (identity not first) and
    'scaled' (element 0 = 2 x identity / sqrt(d), still a real multiple) check the exact host test."""
    rng = np.random.default_rng(1000 + 10*d + n_nops)
    G, n_omega = 53, 211
    _, _, n_opers, n_coeffs, dt, H = _setup(rng, d, G, n_nops)
    n_opers = n_opers + rng.standard_normal(n_nops)[:, None, None]*np.eye(d)   # traces != 0
    ev, V, Q = oracle.diagonalize(H, dt)
    basis = oracle.ggm_basis(d) if btype == 'ggm' else oracle.pauli_basis(1)
    if btype == 'shuffled':
        basis = basis[[1, 0, 2, 3]]
    if btype == 'scaled':
        basis = basis.copy()
        basis[0] *= 2
    omega = np.concatenate(([0.0], np.geomspace(1e-3, 60, n_omega - 1)))
    t = np.concatenate(([0], dt.cumsum()))
    B = engine.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers,
                                                             n_coeffs, dt, t)
    B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs, dt, t)
    k0 = 1 if btype == 'shuffled' else 0
    for j in range(n_nops):
        assert np.abs(B_o[j, k0]).max() > 1e-3*np.abs(B_o[j]).max()    # the identity row matters
        assert nerr(B[j], B_o[j]) < TOL
        assert nerr(B[j, k0], B_o[j, k0]) < TOL


@pytest.mark.parametrize('d,n_nops', [(3, 2), (2, 3), (3, 1), (2, 1)])
def test_control_matrix_special_frequencies(engine, d, n_nops):
    """omega = 0 (the exact-zero branch of numeric.py:162-165), negative omega, omega == -Omega_mn
    (resonance) and tiny omega (series branch), list input, t=None, out=...; (3, 2) runs on the tensor
    path, the other shapes (<= 16 rows) on the DFMA variant."""
    rng = np.random.default_rng(5)
    G = 12
    _, _, n_opers, n_coeffs, dt, H = _setup(rng, d, G, n_nops)
    ev, V, Q = oracle.diagonalize(H, dt)
    basis = oracle.ggm_basis(d)
    res = [ev[3, 1] - ev[3, 0], ev[5, 0] - ev[5, d - 1], ev[0, d - 1] - ev[0, 1]]
    omega = [0.0, 1e-10, -1e-10, 1e-5, -3.3, 2.0, 1e3, -1e4] + res + [-r for r in res]
    B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs, dt)
    B = engine.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers,
                                                             n_coeffs, dt)
    assert nerr(B, B_o) < TOL
    out = np.full_like(B_o, np.nan)
    ret = engine.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers,
                                                               n_coeffs, dt, out=out)
    assert ret is out and nerr(out, B_o) < TOL


@pytest.mark.parametrize('d,G,n_nops,btype', [(3, 1, 2, 'ggm'), (3, 2, 1, 'ggm'), (4, 1, 6, 'pauli'),
                                              (4, 2, 3, 'pauli'), (4, 3, 2, 'ggm'), (5, 3, 2, 'ggm'),
                                              (7, 2, 1, 'ggm'), (8, 1, 3, 'pauli'), (16, 1, 2, 'ggm'),
                                              (16, 2, 3, 'pauli')])
def test_control_matrix_short_pulses(engine, d, G, n_nops, btype):
    """Gate pulses of one to three segments (the constituents of a concatenation, config 5) take the
    transposed operand layout (4 level pairs of one segment per DMMA instead of 4 segments); includes
    omega = 0, negative and resonant frequencies, for which padding slots must stay finite."""
    rng = np.random.default_rng(1000 + 37*d + G)
    _, _, n_opers, n_coeffs, dt, H = _setup(rng, d, G, n_nops)
    ev, V, Q = oracle.diagonalize(H, dt)
    basis = oracle.pauli_basis(int(np.log2(d))) if btype == 'pauli' else oracle.ggm_basis(d)
    res = [ev[0, d - 1] - ev[0, 0], ev[G - 1, 0] - ev[G - 1, d - 1], ev[0, 1] - ev[0, 0]]
    omega = np.concatenate(([0.0, 1e-12, -2.5], res, np.geomspace(1e-3, 80, 70)))
    B = engine.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers,
                                                             n_coeffs, dt)
    B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs, dt)
    assert np.isfinite(B).all()
    for j in range(n_nops):
        assert nerr(B[j], B_o[j]) < TOL


def test_control_matrix_gauge_invariance(engine):
    """Random eigenvector phases must not change the control matrix."""
    rng = np.random.default_rng(11)
    d, G, n_nops = 4, 15, 3
    _, _, n_opers, n_coeffs, dt, H = _setup(rng, d, G, n_nops)
    ev, V, Q = oracle.diagonalize(H, dt)
    basis = oracle.pauli_basis(2)
    omega = np.geomspace(1e-2, 30, 90)
    V2 = V*np.exp(1j*rng.uniform(0, 2*np.pi, (G, 1, d)))
    f = engine.numeric.calculate_control_matrix_from_scratch
    B1 = f(ev, V, Q, omega, basis, n_opers, n_coeffs, dt)
    B2 = f(ev, V2, Q, omega, basis, n_opers, n_coeffs, dt)
    assert nerr(B2, B1) < 1e-12


def test_control_matrix_non_hermitian_operators(engine):
    """Non-Hermitian noise operators and basis elements take the row-expansion path."""
    rng = np.random.default_rng(13)
    d, G = 3, 9
    _, _, _, n_coeffs, dt, H = _setup(rng, d, G, 2)
    n_opers = rng.standard_normal((2, d, d)) + 1j*rng.standard_normal((2, d, d))
    basis = rng.standard_normal((5, d, d)) + 1j*rng.standard_normal((5, d, d))
    ev, V, Q = oracle.diagonalize(H, dt)
    omega = np.geomspace(1e-2, 30, 70)
    f = engine.numeric.calculate_control_matrix_from_scratch
    for no, bs in ((n_opers, oracle.ggm_basis(d)), (rand_herm(rng, d, 2), basis), (n_opers, basis)):
        B = f(ev, V, Q, omega, bs, no, n_coeffs, dt)
        B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega, bs, no, n_coeffs, dt)
        assert nerr(B, B_o) < TOL


def test_control_matrix_large_phase_arguments(engine):
    """|omega t| ~ 1e6 as in config 2 (phase argument needs the host-computed t and a reduction that
    stays accurate for large arguments)."""
    rng = np.random.default_rng(17)
    d, G, n_nops = 2, 400, 3
    c_opers = np.array([[[0, .5], [.5, 0]], [[0, -.5j], [.5j, 0]]])
    n_opers = np.array([[[0, .5], [.5, 0]], [[0, -.5j], [.5j, 0]], [[.5, 0], [0, -.5]]])
    c_coeffs = rng.standard_normal((2, G))*np.pi
    dt = np.full(G, 12.5)   # tau = 5000
    H = oracle.hamiltonian_from_coeffs(c_opers, c_coeffs)
    ev, V, Q = oracle.diagonalize(H, dt)
    omega = np.geomspace(2*np.pi*1e-2/5000, 1257.0, 200)
    assert omega[-1]*dt.sum() > 5e6
    n_coeffs = np.ones((n_nops, G))
    B = engine.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, oracle.pauli_basis(1),
                                                             n_opers, n_coeffs, dt)
    B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega, oracle.pauli_basis(1), n_opers,
                                             n_coeffs, dt)
    assert nerr(B, B_o) < TOL


def test_control_matrix_raises(engine):
    rng = np.random.default_rng(3)
    _, _, n_opers, n_coeffs, dt, H = _setup(rng, 2, 4, 2)
    ev, V, Q = oracle.diagonalize(H, dt)
    f = engine.numeric.calculate_control_matrix_from_scratch
    with pytest.raises(ValueError):
        f(ev, V, Q[:-1], [1.0], oracle.pauli_basis(1), n_opers, n_coeffs, dt)


@pytest.mark.parametrize('n_nops,n_basis,n_omega', [(1, 4, 1), (3, 4, 300), (6, 16, 1001),
                                                     (2, 9, 77), (5, 64, 130)])
@pytest.mark.parametrize('which', ['fidelity', 'generalized'])
def test_filter_function(engine, n_nops, n_basis, n_omega, which):
    rng = np.random.default_rng(n_nops*n_basis + n_omega)
    B = rng.standard_normal((n_nops, n_basis, n_omega)) + 1j*rng.standard_normal(
        (n_nops, n_basis, n_omega))
    if which == 'generalized' and n_basis > 16:
        B = B[:, :16]
    F = engine.numeric.calculate_filter_function(B, which)
    F_o = oracle.filter_function(B, which)
    assert F.shape == F_o.shape
    assert nerr(F, F_o) < 1e-13


@pytest.mark.parametrize('P,n_nops,n_basis,n_omega', [
    (1, 18, 256, 100),   # config 5's shape: one 18-row panel, 6 x 6 tiles
    (1, 4, 32, 31), (1, 8, 36, 64), (1, 9, 64, 65), (1, 12, 37, 70), (1, 13, 35, 33), (1, 16, 64, 96),
    (1, 17, 33, 45), (1, 19, 34, 40), (1, 32, 32, 33), (1, 40, 35, 10),   # two and three 16-row panels
    (3, 7, 64, 50), (2, 9, 49, 1)])
def test_filter_function_gram(engine, P, n_nops, n_basis, n_omega):
    """The one-pass Gram kernel (n_basis >= 32, at least 4 rows): every tile shape, panel pairs, padding
    rows, basis sums that are not a multiple of the stage depth, frequency tails; exact Hermiticity."""
    rng = np.random.default_rng(P*1000 + n_nops*n_basis + n_omega)
    shape = (P, n_nops, n_basis, n_omega)
    B = rng.standard_normal(shape) + 1j*rng.standard_normal(shape)
    if P == 1:
        F = engine.numeric.calculate_filter_function(B[0], 'fidelity')
        F_o = oracle.filter_function(B[0], 'fidelity')
        assert np.array_equal(F, F.conj().swapaxes(0, 1))
        assert np.all(F[np.arange(n_nops), np.arange(n_nops)].imag == 0)
    else:
        F = engine.numeric.calculate_pulse_correlation_filter_function(B, 'fidelity')
        F_o = oracle.pulse_correlation_filter_function(B, 'fidelity')
    assert F.shape == F_o.shape
    assert nerr(F, F_o) < 1e-13


@pytest.mark.parametrize('which', ['fidelity', 'generalized'])
def test_pulse_correlation_filter_function(engine, which):
    rng = np.random.default_rng(23)
    B = rng.standard_normal((3, 2, 9, 55)) + 1j*rng.standard_normal((3, 2, 9, 55))
    F = engine.numeric.calculate_pulse_correlation_filter_function(B, which)
    F_o = oracle.pulse_correlation_filter_function(B, which)
    assert F.shape == F_o.shape and nerr(F, F_o) < 1e-13
    with pytest.raises(ValueError):
        engine.numeric.calculate_pulse_correlation_filter_function(B[0], which)
    with pytest.raises(ValueError):
        engine.numeric.calculate_filter_function(B[0], 'foo')


@pytest.mark.parametrize('P,n_nops,n_basis,n_omega', [(2, 1, 4, 301), (5, 3, 16, 200),
                                                      (3, 2, 9, 64), (4, 2, 64, 100), (1, 2, 4, 10),
                                                      (3, 2, 36, 77), (2, 3, 49, 9), (3, 1, 256, 41),
                                                      (1, 2, 64, 17), (6, 1, 100, 33)])
@pytest.mark.parametrize('which', ['total', 'correlations'])
def test_control_matrix_from_atomic(engine, P, n_nops, n_basis, n_omega, which):
    rng = np.random.default_rng(P*n_basis + n_omega)
    shape = (P, n_nops, n_basis, n_omega)
    atomic = rng.standard_normal(shape) + 1j*rng.standard_normal(shape)
    phases = oracle.cexp(rng.standard_normal((P - 1, n_omega))*10)
    Q = rng.standard_normal((P - 1, n_basis, n_basis))
    out = engine.numeric.calculate_control_matrix_from_atomic(phases, atomic, Q, which=which)
    out_o = oracle.control_matrix_from_atomic(phases, atomic, Q, which)
    assert out.shape == out_o.shape and nerr(out, out_o) < 1e-13
    Qc = Q + 1j*rng.standard_normal(Q.shape)
    out = engine.numeric.calculate_control_matrix_from_atomic(phases, atomic, Qc, which=which)
    assert nerr(out, oracle.control_matrix_from_atomic(phases, atomic, Qc, which)) < 1e-13


def test_from_atomic_memory_layout(engine):
    """C / F / non-contiguous input => same flags out (reference tests/test_sequencing.py:65-93)."""
    rng = np.random.default_rng(29)
    P, n_nops, n_basis, n_omega = 3, 2, 4, 30
    atomic = rng.standard_normal((P, n_nops, n_basis, n_omega)) + 0j
    phases = oracle.cexp(rng.standard_normal((P - 1, n_omega)))
    Q = rng.standard_normal((P - 1, n_basis, n_basis))
    f = engine.numeric.calculate_control_matrix_from_atomic
    ref = oracle.control_matrix_from_atomic(phases, atomic, Q)
    for which in ('total', 'correlations'):
        ref = oracle.control_matrix_from_atomic(phases, atomic, Q, which)
        out_c = f(phases, np.ascontiguousarray(atomic), Q, which=which)
        out_f = f(phases, np.asfortranarray(atomic), Q, which=which)
        other = np.ascontiguousarray(atomic.swapaxes(-1, -2)).swapaxes(-1, -2)
        out_n = f(phases, other, Q, which=which)
        assert out_c.flags.c_contiguous
        assert out_f.flags.f_contiguous
        assert not out_n.flags.c_contiguous and not out_n.flags.f_contiguous
        for o in (out_c, out_f, out_n):
            assert nerr(o, ref) < 1e-13


@pytest.mark.parametrize('n_omega', [501, 2049, 2050, 5000, 70001])
def test_infidelity_integral(engine, n_omega):
    """501: one block per output; >= 2049 intervals: chunked over several blocks per output with the
    partial sums added in chunk order (2049 / 2050 straddle the switch, 70001 hits the chunk cap);
    repeated calls reuse the ticket counters."""
    rng = np.random.default_rng(31)
    n_nops, d = 4, 2
    B = rng.standard_normal((n_nops, 4, n_omega)) + 1j*rng.standard_normal((n_nops, 4, n_omega))
    F = oracle.filter_function(B)
    omega = np.sort(rng.random(n_omega))*10
    S1 = 1/(1 + omega**2)
    idx = np.array([2, 0, 3])
    S2 = np.array([S1*(i + 1) for i in range(3)])
    S3 = np.einsum('a,b,o->abo', [1, 2, 3], [1, 2, 3], S1) + 0j
    S3[0, 1] += 1j*omega
    S3[1, 0] -= 1j*omega
    f = engine.numeric._integrate_against_spectrum
    for S in (S1, S2, S3, S1):
        got = f(F, S, omega, idx, d)
        want = oracle.infidelity_from_filter_function(F, S, omega, d, idx)
        assert got.shape == want.shape and nerr(got, want) < 1e-13
        assert np.array_equal(got, f(F, S, omega, idx, d))     # deterministic
    # leading pulse-correlation axes
    Fpc = oracle.pulse_correlation_filter_function(np.stack([B, 2*B]))
    got = f(Fpc, S2, omega, idx, d)
    assert nerr(got, oracle.infidelity_from_filter_function(Fpc, S2, omega, d, idx)) < 1e-13
    with pytest.raises(ValueError):
        f(F, S2[:2], omega, idx, d)
    bad = S3.copy()
    bad[0, 1] += 1.0
    with pytest.raises(ValueError):
        f(F, bad, omega, idx, d)


@pytest.mark.parametrize('d,btype', [(2, 'pauli'), (3, 'ggm'), (4, 'pauli'), (8, 'ggm')])
def test_liouville_representation(engine, d, btype):
    rng = np.random.default_rng(d)
    H = rand_herm(rng, d, 3)
    w, v = np.linalg.eigh(H)
    U = (v*np.exp(-1j*w)[:, None, :]) @ v.conj().transpose(0, 2, 1)
    basis = engine.Basis.pauli(int(np.log2(d))) if btype == 'pauli' else engine.Basis.ggm(d)
    L = engine.liouville_representation(U, basis)
    L_o = oracle.liouville_representation(U, np.asarray(basis))
    assert L.dtype == np.float64 and nerr(L, L_o) < 1e-13
    # real orthogonal (reference tests/test_superoperator.py:35-74)
    assert nerr(L @ L.transpose(0, 2, 1), np.broadcast_to(np.eye(d*d), L.shape)) < 1e-12
    assert engine.liouville_representation(U[0], basis).shape == (d*d, d*d)


def test_cexp(engine):
    x = np.random.default_rng(1).standard_normal((7, 13))*1e3
    assert nerr(engine.util.cexp(x), np.exp(1j*x)) < 1e-15


@pytest.mark.parametrize('d,G,n_nops,btype,n_omega', [(2, 7, 3, 'pauli', 150), (3, 5, 2, 'ggm', 33),
                                                      (4, 1, 2, 'pauli', 64), (6, 4, 1, 'ggm', 17)])
def test_control_matrix_intermediates(engine, d, G, n_nops, btype, n_omega):
    """cache_intermediates=True (reference numeric.py:828-879, tests/test_core.py:604-642): the control
    matrix and every entry of the intermediates dict, for fixed input eigenvectors (so nothing is
    gauge dependent); includes omega = 0 (exact-zero branch of the integral) and negative omega."""
    rng = np.random.default_rng(1000*d + G)
    c_opers, c_coeffs, n_opers, n_coeffs, dt, H = _setup(rng, d, G, n_nops)
    ev, V, Q = oracle.diagonalize(H, dt)
    basis = oracle.pauli_basis(int(np.log2(d))) if btype == 'pauli' else oracle.ggm_basis(d)
    omega = np.concatenate(([0.0, -0.3], np.geomspace(1e-3, 30, n_omega - 2)))
    B, inter = engine.numeric.calculate_control_matrix_from_scratch(
        ev, V, Q, omega, basis, n_opers, n_coeffs, dt, cache_intermediates=True)
    B_ref, inter_ref = oracle.control_matrix_intermediates(ev, V, Q, omega, basis, n_opers, n_coeffs,
                                                           dt)
    assert nerr(B, B_ref) < TOL
    assert sorted(inter) == sorted(inter_ref)
    for key, ref in inter_ref.items():
        assert inter[key].shape == ref.shape, key
        assert inter[key].dtype == np.complex128
        if ref.size:
            assert nerr(inter[key], ref) < TOL, key
    # the fused kernel and the materialising kernels agree
    fused = engine.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers,
                                                                 n_coeffs, dt)
    assert nerr(fused, B) < 1e-12
    out = np.empty_like(B)
    res, _ = engine.numeric.calculate_control_matrix_from_scratch(
        ev, V, Q, omega, basis, n_opers, n_coeffs, dt, cache_intermediates=True, out=out)
    assert res is out and nerr(out, B) == 0


def test_control_matrix_intermediates_long_pulse(engine):
    """More segments than one grid dimension holds (G > 65535): the materialising kernels walk the
    segment axis in slices; checked against the fused kernel (whole result) and the oracle (tail slice
    of the per-segment arrays, which only exists if the second slice was written)."""
    d, G, n_nops, n_omega = 2, 66000, 1, 3
    rng = np.random.default_rng(7)
    c_opers, c_coeffs, n_opers, n_coeffs, dt, H = _setup(rng, d, G, n_nops)
    ev, V, Q = engine.numeric.diagonalize(H, dt)
    basis = oracle.pauli_basis(1)
    omega = np.array([0.0, 0.37, 2.1])
    B, inter = engine.numeric.calculate_control_matrix_from_scratch(
        ev, V, Q, omega, basis, n_opers, n_coeffs, dt, cache_intermediates=True)
    fused = engine.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers,
                                                                 n_coeffs, dt)
    assert nerr(B, fused) < TOL
    assert inter['control_matrix_step'].shape == (G, n_nops, 4, n_omega)
    tail = slice(G - 40, G)
    t = np.concatenate(([0.0], dt.cumsum()))
    _, ref = oracle.control_matrix_intermediates(ev[tail], V[tail], Q[G - 40:], omega, basis, n_opers,
                                                 n_coeffs[:, tail], dt[tail], t=t[G - 40:])
    for key in ('phase_factors', 'first_order_integral', 'control_matrix_step'):
        assert nerr(inter[key][tail], ref[key]) < TOL, key


@pytest.mark.parametrize('d,G,n_nops,n_omega', [(8, 40, 3, 70), (8, 23, 1, 33), (8, 9, 6, 40), (16, 2, 18, 64),
                                                (16, 1, 3, 40), (4, 50, 6, 70),
                                                # 10, 5 and 7 row tiles (no whole padding tile), 4 x 10 tiles
                                                (4, 44, 5, 70), (3, 30, 4, 50), (3, 21, 6, 40), (8, 6, 5, 33)])
def test_static_and_generic_tensor_kernels_agree(engine, d, G, n_nops, n_omega):
    """The statically scheduled DMMA kernel (d = 4; d = 8 with the pass in several pieces; the transposed
    d = 16 layout of short pulses) against the generic kernel (FFB_CTRLMAT_STATIC=0, read per call) and
    the oracle.  Uniform and non-uniform dt, omega = 0 and a negative frequency."""
    import os
    rng = np.random.default_rng(31*d + G + n_nops)
    c_opers, c_coeffs, n_opers, n_coeffs, dt, H = _setup(rng, d, G, n_nops)
    if G % 2:
        dt = np.full(G, 0.37)
    ev, V, Q = oracle.diagonalize(H, dt)
    basis = oracle.ggm_basis(d)
    omega = np.concatenate(([0.0, -0.21], np.geomspace(1e-3, 25, n_omega - 2)))
    f = engine.numeric.calculate_control_matrix_from_scratch
    old = os.environ.get('FFB_CTRLMAT_STATIC')
    try:
        os.environ['FFB_CTRLMAT_STATIC'] = '1'
        B_static = f(ev, V, Q, omega, basis, n_opers, n_coeffs, dt)
        os.environ['FFB_CTRLMAT_STATIC'] = '0'
        B_generic = f(ev, V, Q, omega, basis, n_opers, n_coeffs, dt)
    finally:
        if old is None:
            os.environ.pop('FFB_CTRLMAT_STATIC', None)
        else:
            os.environ['FFB_CTRLMAT_STATIC'] = old
    B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs, dt)
    for j in range(n_nops):
        assert nerr(B_static[j], B_o[j]) < 1e-12
        assert nerr(B_generic[j], B_o[j]) < 1e-12
        assert nerr(B_static[j], B_generic[j]) < 1e-13
