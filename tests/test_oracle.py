"""The CPU oracle against the reference's own golden vectors and fixtures (no GPU needed).

Pins ``oracle/ff_oracle.py``:
  * seeded infidelities of reference tests/test_precision.py:495-551 (hard-coded there, atol 1e-12),
  * README value 0.00253303 (README.md:57-60),
  * analytic dynamical-decoupling filter functions (analytic.py:59-88 via tests/test_precision.py:75-182),
  * fixtures generated from the unmodified reference by oracle/gen_golden.py.
"""
import os

import numpy as np
import pytest

import ff_oracle as oracle
from helpers import dd_hamiltonian, nerr, rand_pulse_arrays

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# values hard-coded in reference tests/test_precision.py:510-529
REF_INFIDS = (
    [2.1571674053883583, 2.1235628100639845],
    [1.7951695420688032, 2.919850951578396],
    [0.4327173760925169, 0.817672660809546],
    [2.1571674053883583, 2.919850951578396],
    [[1.7951695420688032, -1.1595479985471822], [-1.1595479985471822, 2.919850951578396]],
    [0.8247284959004152, 2.495561429509174],
    [0.854760904366362, 3.781670732974073],
    [0.24181791977082442, 1.122626106375816],
    [0.8247284959004152, 3.781670732974073],
    [[0.854760904366362, -0.16574972846239408], [-0.16574972846239408, 3.781670732974073]],
    [2.9464977186365267, 0.8622319594213088],
    [2.8391133843027525, 0.678843575761492],
    [0.813728718501677, 0.16950739577216872],
    [2.9464977186365267, 0.678843575761492],
    [[2.8391133843027525, 0.2725782717379744], [0.2725782717379744, 0.678843575761492]],
)

SPECTRA = [
    lambda S0, omega: S0*abs(omega)**0,
    lambda S0, omega: S0/abs(omega)**0.7,
    lambda S0, omega: S0*np.exp(-abs(omega)),
    lambda S0, omega: np.array([S0*abs(omega)**0, S0/abs(omega)**0.7]),
    lambda S0, omega: np.array([[S0/abs(omega)**0.7, S0/(1 + omega**2) + 1j*S0*omega],
                                [S0/(1 + omega**2) - 1j*S0*omega, S0/abs(omega)**0.7]]),
]


def oracle_filter_function(c_opers, c_coeffs, n_opers, n_coeffs, dt, basis, omega):
    H = oracle.hamiltonian_from_coeffs(c_opers, c_coeffs)
    ev, V, Q = oracle.diagonalize(H, dt)
    B = oracle.control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs, dt)
    return B, oracle.filter_function(B)


def seeded_pulses():
    """Same draw order as the reference test (rand_pulse_sequence, then S0)."""
    rng = np.random.default_rng(seed=123456789)
    for d in (2, 3, 4):
        c_opers, c_coeffs, c_ids, n_opers, n_coeffs, n_ids, dt = rand_pulse_arrays(rng, d, 10, 2, 3)
        S0 = np.abs(rng.standard_normal())
        # PulseSequence sorts operators by identifier (pulse_sequence.py:1333-1337)
        co, no = np.argsort(c_ids), np.argsort(n_ids)
        yield d, c_opers[co], c_coeffs[co], n_opers[no], n_coeffs[no], dt, S0


def test_seeded_infidelities_hardcoded():
    omega = np.geomspace(0.1, 10, 51)
    count = 0
    for d, c_opers, c_coeffs, n_opers, n_coeffs, dt, S0 in seeded_pulses():
        _, F = oracle_filter_function(c_opers, c_coeffs, n_opers, n_coeffs, dt,
                                      oracle.ggm_basis(d), omega)
        for spec in SPECTRA:
            infid = oracle.infidelity_from_filter_function(F, spec(S0, omega), omega, d,
                                                           idx=np.array([0, 1]))
            np.testing.assert_allclose(infid, REF_INFIDS[count], atol=1e-12, rtol=0)
            count += 1
    assert count == 15


def test_seeded_fixture_matches_hardcoded_and_oracle():
    g = np.load(os.path.join(GOLDEN, 'infidelity_seeded.npz'))
    count = 0
    for d in (2, 3, 4):
        B, F = oracle_filter_function(g[f'd{d}_c_opers'], g[f'd{d}_c_coeffs'], g[f'd{d}_n_opers'],
                                      g[f'd{d}_n_coeffs'], g[f'd{d}_dt'], g[f'd{d}_basis'],
                                      g[f'd{d}_omega'])
        assert nerr(B, g[f'd{d}_control_matrix']) < 1e-12
        assert nerr(F, g[f'd{d}_filter_function']) < 1e-12
        for i in range(5):
            np.testing.assert_allclose(g[f'd{d}_infid{i}'], REF_INFIDS[count], atol=1e-12, rtol=0)
            count += 1


def test_readme_value():
    X, Y, Z = oracle.paulis[1:]
    dt = np.array([1.0, 1.0])
    omega = oracle.sample_frequencies(dt.sum(), dt.min())
    _, F = oracle_filter_function(np.array([X/2, Y/2]), np.array([[0, np.pi], [np.pi/2, 0]]),
                                  np.array([Z/2]), np.ones((1, 2)), dt, oracle.ggm_basis(2), omega)
    infid = oracle.infidelity_from_filter_function(F, 1e-2/omega, omega, 2)
    assert abs(infid[0] - 0.00253303) < 5e-9


def analytic_dd(kind, z, n):
    """Closed forms of reference analytic.py:59-88 (restated)."""
    if kind == 'fid':
        return 2*np.sin(z/2)**2
    if kind == 'se':
        return 8*np.sin(z/4)**4
    if kind == 'pdd':
        trig = np.cos(z/2) if n % 2 == 0 else np.sin(z/2)
        return 2*np.tan(z/(2*n + 2))**2*trig**2
    if kind == 'cpmg':
        trig = np.sin(z/2) if n % 2 == 0 else np.cos(z/2)
        return 8*np.sin(z/4/n)**4*trig**2/np.cos(z/2/n)**2
    if kind == 'cdd':
        return 2**(2*n + 1)*np.sin(z/2**(n + 1))**2*np.prod(
            [np.sin(z/2**(k + 1))**2 for k in range(1, n + 1)], axis=0)
    if kind == 'udd':
        return np.abs(np.sum([(-1)**k*np.exp(1j*z/2*np.cos(np.pi*k/(n + 1)))
                              for k in range(-n - 1, n + 1)], axis=0))**2/2
    raise ValueError(kind)


DD_CASES = [('cpmg', 'se', 1, 1e-8), ('cpmg', 'cpmg', 6, 1e-9), ('udd', 'udd', 6, 1e-9),
            ('pdd', 'pdd', 6, 1e-9), ('cdd', 'cdd', 3, 1e-9)]


@pytest.mark.parametrize('dd_type,formula,n,tau_pi', DD_CASES)
def test_analytic_dd(dd_type, formula, n, tau_pi):
    tau = np.pi
    H_c, dt = dd_hamiltonian(n, tau=tau, tau_pi=tau_pi, dd_type=dd_type)
    omega = np.logspace(0, 3, 100)
    omega = np.concatenate([-omega[::-1], omega])
    Z = oracle.paulis[3]
    _, F = oracle_filter_function(np.array([H_c[0][0]]), np.array([H_c[0][1]]), np.array([Z/2]),
                                  np.ones((1, len(dt))), dt, oracle.ggm_basis(2), omega)
    # tolerances of the reference's assertArrayAlmostEqual(..., atol=1e-10): rtol defaults to 1e-7
    np.testing.assert_allclose(F[0, 0].real*omega**2, analytic_dd(formula, omega*tau, n),
                               atol=1e-10, rtol=1e-7)


def test_random_pulse_fixture():
    g = np.load(os.path.join(GOLDEN, 'random_pulses.npz'))
    for d in (2, 3, 4, 5):
        t = f'd{d}'
        H = oracle.hamiltonian_from_coeffs(g[f'{t}_c_opers'], g[f'{t}_c_coeffs'])
        ev, V, Q = oracle.diagonalize(H, g[f'{t}_dt'])
        assert nerr(ev, g[f'{t}_eigvals']) < 1e-12 and nerr(Q, g[f'{t}_propagators']) < 1e-12
        B = oracle.control_matrix_from_scratch(ev, V, Q, g[f'{t}_omega'], g[f'{t}_basis'],
                                               g[f'{t}_n_opers'], g[f'{t}_n_coeffs'], g[f'{t}_dt'])
        assert nerr(B, g[f'{t}_control_matrix']) < 1e-12
        assert nerr(oracle.filter_function(B), g[f'{t}_filter_function']) < 1e-12
        if d <= 3:
            assert nerr(oracle.filter_function(B, 'generalized'),
                        g[f'{t}_filter_function_gen']) < 1e-12
        assert nerr(oracle.total_phases(g[f'{t}_omega'], g[f'{t}_dt'].sum()),
                    g[f'{t}_total_phases']) < 1e-12
        assert nerr(oracle.liouville_representation(Q[-1], g[f'{t}_basis']),
                    g[f'{t}_total_propagator_liouville']) < 1e-12
        om = g[f'{t}_int_omega']
        Bi = oracle.control_matrix_from_scratch(ev, V, Q, om, g[f'{t}_basis'], g[f'{t}_n_opers'],
                                                g[f'{t}_n_coeffs'], g[f'{t}_dt'])
        infid = oracle.infidelity_from_filter_function(oracle.filter_function(Bi),
                                                       g[f'{t}_int_spectrum'], om, d)
        assert nerr(infid, g[f'{t}_infidelity']) < 1e-12


def test_concatenation_fixture():
    """from_atomic over the fixture's constituents reproduces the reference's concatenate()."""
    g = np.load(os.path.join(GOLDEN, 'concatenation.npz'))
    omega = g['omega']
    basis = g['p0_basis']
    n_ids = list(g['n_ids'])
    Z, Xn = oracle.paulis[3]/2, oracle.paulis[1]/2
    all_opers = {'Z': Z, 'Xn': Xn}
    const = {'Z': 1.0, 'Xn': 0.5}
    atomic, phases, Ls = [], [], []
    for i in range(4):
        dt = g[f'p{i}_dt']
        H = oracle.hamiltonian_from_coeffs(g[f'p{i}_c_opers'], g[f'p{i}_c_coeffs'])
        ev, V, Q = oracle.diagonalize(H, dt)
        n_opers = np.array([all_opers[k] for k in n_ids])
        n_coeffs = np.array([np.full(len(dt), const[k]) for k in n_ids])
        atomic.append(oracle.control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers,
                                                         n_coeffs, dt))
        phases.append(oracle.total_phases(omega, dt.sum()))
        Ls.append(oracle.liouville_representation(Q[-1], basis))
    cum_phase = np.cumprod(phases[:-1], axis=0)
    cum_L = [Ls[0]]
    for L in Ls[1:-1]:
        cum_L.append(L @ cum_L[-1])
    pc = oracle.control_matrix_from_atomic(cum_phase, np.array(atomic), np.array(cum_L),
                                           'correlations')
    assert nerr(pc, g['control_matrix_pc']) < 1e-12
    assert nerr(pc.sum(0), g['control_matrix']) < 1e-12
    assert nerr(g['control_matrix_scratch'], g['control_matrix']) < 1e-12
    Fpc = oracle.pulse_correlation_filter_function(pc)
    assert nerr(Fpc, g['filter_function_pc']) < 1e-12
    assert nerr(Fpc.sum((0, 1)), g['filter_function']) < 1e-12
    S = 1e-2/omega
    assert nerr(oracle.infidelity_from_filter_function(Fpc, S, omega, 2), g['infidelity_pc']) < 1e-12


def test_workload_generators_pinned():
    """The bench workloads (reduced size) through the oracle equal the reference's results."""
    import workloads
    g = np.load(os.path.join(GOLDEN, 'workloads_small.npz'))
    for name, kwargs in (('c2', dict(G=64, n_omega=96)), ('c3', dict(G=40, n_omega=64))):
        wl = workloads.get(name, **kwargs)
        order = np.argsort(wl.n_ids)
        assert list(np.asarray(wl.n_ids)[order]) == list(g[f'{name}_n_ids'])
        B, F = oracle_filter_function(wl.c_opers, wl.c_coeffs, wl.n_opers[order],
                                      wl.n_coeffs[order], wl.dt, wl.basis, wl.omega)
        assert nerr(B, g[f'{name}_control_matrix']) < 1e-12
        assert nerr(F, g[f'{name}_filter_function']) < 1e-12
        infid = oracle.infidelity_from_filter_function(F, wl.spectrum, wl.omega, wl.d)
        assert nerr(infid, g[f'{name}_infidelity']) < 1e-12
    wl = workloads.get('c1')
    assert abs(g['c1_infidelity'][0] - wl.extra['expected_infidelity']) < 5e-9


def test_elementwise_helpers():
    x = np.random.default_rng(0).standard_normal(500)*50
    np.testing.assert_allclose(oracle.cexp(x), np.exp(1j*x), atol=1e-15)
    np.testing.assert_allclose(oracle.cexpm1(x), np.expm1(1j*x), atol=1e-14)
    from scipy import integrate
    xs = np.sort(np.random.default_rng(1).random(100))
    f = np.random.default_rng(2).standard_normal((4, 100))
    np.testing.assert_allclose(oracle.integrate(f, xs), integrate.trapezoid(f, xs), atol=1e-15)


def test_first_order_integral_against_quadrature():
    """Reference tests/test_precision.py:469-493: the closed-form integral vs 1001-point trapezoid,
    including omega = 0 and 1e-10."""
    rng = np.random.default_rng(5)
    d, dt = 3, 0.7
    E = np.array([0.0, 1e-10, 0.5, 3.0])
    eigvals = np.sort(rng.standard_normal(d))
    t = np.linspace(0, dt, 1001)
    dE = np.subtract.outer(eigvals, eigvals)
    numeric_int = np.trapezoid(np.exp(1j*np.multiply.outer(np.add.outer(E, dE), t)), t)
    np.testing.assert_allclose(oracle.first_order_integral(E, eigvals, dt), numeric_int, atol=1e-4)


def test_cnot_fixture():
    """The reference's experimental fixture examples/data/CNOT.mat (exchange-coupled singlet-triplet
    CNOT, 250 segments, 6-dimensional subspace, padded two-qubit Pauli basis of 15 elements), run
    through the reference by oracle/gen_golden.py: the oracle reproduces the reference's propagator,
    control matrix, filter function and infidelities (tests/test_precision.py:184-216, :274-311)."""
    g = np.load(os.path.join(GOLDEN, 'cnot.npz'))
    omega = g['cnot_omega']
    H = oracle.hamiltonian_from_coeffs(g['cnot_c_opers'], g['cnot_c_coeffs'])
    ev, V, Q = oracle.diagonalize(H, g['cnot_dt'])
    assert nerr(ev, g['cnot_eigvals']) < 1e-13
    assert nerr(Q[-1], g['cnot_total_propagator']) < 1e-12
    B = oracle.control_matrix_from_scratch(ev, V, Q, omega, g['cnot_basis'], g['cnot_n_opers'],
                                           g['cnot_n_coeffs'], g['cnot_dt'])
    assert nerr(B, g['cnot_control_matrix']) < 1e-12
    F = oracle.filter_function(B)
    assert nerr(F, g['cnot_filter_function']) < 1e-12
    cnot = np.zeros((4, 4))
    cnot[0, 0] = cnot[1, 1] = cnot[2, 3] = cnot[3, 2] = 1
    U = Q[-1][1:5, 1:5]
    phase = np.trace(cnot.T @ U)
    assert np.abs(U - cnot*phase/np.abs(phase)).max() < 1e-4   # CNOT up to a global phase
    assert 1 - np.abs(phase)/4 < 1e-5                          # (optimised gate: 1.4e-5 off)
    ids = [str(k) for k in g['cnot_n_ids']]      # PulseSequence sorts operators by identifier
    eps = np.array([ids.index(k) for k in ('eps_12', 'eps_23', 'eps_34')])
    for i, alpha in enumerate((0.0, 0.7)):
        S = g['cnot_A'][i]/omega**alpha
        infid = oracle.infidelity_from_filter_function(F, S, omega, 4, idx=eps)
        np.testing.assert_allclose(infid, g[f'cnot_infid_{i}'], rtol=1e-10)
        assert abs(1 - infid.sum()/g['cnot_infid_MC'][i]) < 0.10   # Monte Carlo, as the reference
