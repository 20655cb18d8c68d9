"""Host-side logic of the drop-in shell that needs no GPU: constructor validation, identifier
handling, Hamiltonian concatenation, cache bookkeeping, basis constructors, argument parsing and the
frequency sharding used by the multi-GPU path.  Mirrors reference tests/test_core.py:42-221,
:1124-1232 and tests/test_sequencing.py:507-606 where the behaviour is host-only."""
import numpy as np
import pytest

import ff_oracle as oracle
import filter_functions_b200 as ff
from filter_functions_b200 import distributed as ffd
from filter_functions_b200 import util
from filter_functions_b200.pulse_sequence import concatenate_without_filter_function

X, Y, Z = util.paulis[1:]


def simple_pulse(n_dt=3, ids=('X', 'Y'), n_ids=('Z',)):
    H_c = [[X/2, np.arange(n_dt) + 1.0, ids[0]], [Y/2, np.ones(n_dt), ids[1]]]
    H_n = [[Z/2, np.ones(n_dt), n_ids[0]]]
    return ff.PulseSequence(H_c, H_n, np.full(n_dt, 0.5))


def test_constructor_validation():
    dt = [1.0, 2.0]
    H_c = [[X/2, [1, 2]]]
    H_n = [[Z/2, [1, 1]]]
    with pytest.raises(TypeError):
        ff.PulseSequence(H_c, H_n, 1.0)
    with pytest.raises(ValueError):
        ff.PulseSequence(H_c, H_n, [1.0, -1.0])
    with pytest.raises(ValueError):
        ff.PulseSequence(H_c, H_n, [1.0, 1j])
    with pytest.raises(TypeError):
        ff.PulseSequence(1, H_n, dt)
    with pytest.raises(TypeError):
        ff.PulseSequence([1], H_n, dt)
    with pytest.raises(TypeError):
        ff.PulseSequence([[X/2, 3]], H_n, dt)
    with pytest.raises(ValueError):
        ff.PulseSequence([[X/2, [1, 2, 3]]], H_n, dt)
    with pytest.raises(ValueError):
        ff.PulseSequence([[X/2, [1, 2], 'a'], [Y/2, [1, 2], 'a']], H_n, dt)
    with pytest.raises(ValueError):
        ff.PulseSequence(H_c, [[np.eye(3), [1, 1]]], dt)
    with pytest.raises(ValueError):
        ff.PulseSequence([[np.ones((2, 3)), [1, 2]]], H_n, dt)
    with pytest.raises(ValueError):
        ff.PulseSequence(H_c, H_n, dt, basis=np.eye(2))
    with pytest.raises(ValueError):
        ff.PulseSequence(H_c, H_n, dt, basis=ff.Basis.pauli(2))
    with pytest.raises(TypeError):
        ff.PulseSequence([[object(), [1, 2]]], H_n, dt)


def test_identifiers_and_sorting():
    pulse = ff.PulseSequence([[Y/2, [1, 2], 'b'], [X/2, [3, 4], 'a']], [[Z/2, [1, 1]], [X/2, [2, 2]]],
                             [1, 1])
    assert list(pulse.c_oper_identifiers) == ['a', 'b']
    np.testing.assert_array_equal(pulse.c_opers[0], X/2)
    np.testing.assert_array_equal(pulse.c_coeffs, [[3, 4], [1, 2]])
    assert list(pulse.n_oper_identifiers) == ['B_0', 'B_1']
    assert pulse.d == 2 and len(pulse) == 2 and pulse.basis.btype == 'GGM'
    np.testing.assert_allclose(pulse.t, [0, 1, 2])
    assert pulse.tau == 2 and pulse.duration == 2
    none_ids = ff.PulseSequence([[X/2, [1, 2], None]], [[Z/2, [1, 1], None]], [1, 1])
    assert list(none_ids.c_oper_identifiers) == ['A_0'] and list(none_ids.n_oper_identifiers) == ['B_0']


def test_from_arrays_validation():
    p = simple_pulse()
    args = dict(c_opers=p.c_opers, c_oper_identifiers=p.c_oper_identifiers, c_coeffs=p.c_coeffs,
                n_opers=p.n_opers, n_oper_identifiers=p.n_oper_identifiers, n_coeffs=p.n_coeffs,
                dt=p.dt, basis=p.basis)
    q = ff.PulseSequence.from_arrays(**args)
    assert q == p
    for key, bad in (('c_coeffs', p.c_coeffs[:1]), ('n_oper_identifiers', ['a', 'b']),
                     ('dt', p.dt[:2]), ('n_opers', np.zeros((1, 3, 3))),
                     ('basis', ff.Basis.pauli(2))):
        with pytest.raises(ValueError):
            ff.PulseSequence.from_arrays(**{**args, key: bad})


def test_equality_and_slicing():
    p = simple_pulse()
    assert p == simple_pulse()
    assert not (p == simple_pulse(ids=('X', 'W')))
    assert (p == 1) is False
    # two equal consecutive segments compare equal to the merged one (pulse_sequence.py:386-390)
    a = ff.PulseSequence([[X/2, [1, 1]]], [[Z/2, [1, 1]]], [1, 2])
    b = ff.PulseSequence([[X/2, [1]]], [[Z/2, [1]]], [3])
    assert a == b
    s = p[1:]
    assert len(s) == 2 and np.array_equal(s.c_coeffs, p.c_coeffs[:, 1:])
    assert len(p[0]) == 1
    with pytest.raises(IndexError):
        p[3:]


def test_cache_bookkeeping_without_compute():
    p = simple_pulse()
    assert not p.is_cached('eigvals') and not p.is_cached('control matrix')
    p.omega = [1.0, 2.0]
    assert p.is_cached('frequencies') and isinstance(p.omega, np.ndarray)
    p._frequency_data['control_matrix'] = np.zeros((1, 4, 2), dtype=complex)
    p._data['eigvals'] = np.zeros((3, 2))
    assert p.is_cached('Control Matrix') and p.is_cached('control_matrix')
    assert p.is_cached('eigenvalues')
    p.omega = [1.0, 2.0]                      # same frequencies: keeps the cache
    assert p.is_cached('control_matrix')
    p.omega = np.array([1.0, 3.0])            # different: drops frequency-dependent data
    assert not p.is_cached('control_matrix') and p.is_cached('eigvals')
    assert p.nbytes > 0
    p.cleanup('conservative')
    assert not p.is_cached('eigvals') and p.is_cached('omega')
    p.cleanup('all')
    assert not p.is_cached('omega')
    with pytest.raises(ValueError):
        p.cleanup('everything')
    with pytest.raises(util.CalculationError):
        p.get_pulse_correlation_filter_function()
    with pytest.raises(util.CalculationError):
        p.get_pulse_correlation_control_matrix()
    with pytest.raises(NotImplementedError):
        p.get_filter_function([1.0], order=2)
    with pytest.raises(ValueError):
        p.get_filter_function([1.0], which='foo')
    q = p.__copy__()
    q._data['x'] = 1
    assert 'x' not in p.data
    with pytest.raises(TypeError):
        p @ 1
    with pytest.raises(TypeError):
        p.data['x'] = 1


def test_concatenate_hamiltonians():
    """Operator merging, identifier clashes and inferred noise coefficients, without any numerics
    (reference pulse_sequence.py:1340-1483, tests/test_sequencing.py:507-606)."""
    a = ff.PulseSequence([[X/2, [1, 2], 'X']], [[Z/2, [1, 1], 'Z']], [1, 1])
    b = ff.PulseSequence([[Y/2, [3], 'Y']], [[Z/2, [1], 'Z'], [X/2, [2], 'Xn']], [0.5])
    c, cmap, nmap = concatenate_without_filter_function([a, b], return_identifier_mappings=True)
    assert list(c.c_oper_identifiers) == ['X', 'Y'] and list(c.n_oper_identifiers) == ['Xn', 'Z']
    np.testing.assert_array_equal(c.c_coeffs, [[1, 2, 0], [0, 0, 3]])
    np.testing.assert_array_equal(c.n_coeffs, [[2, 2, 2], [1, 1, 1]])   # constant inferred
    np.testing.assert_array_equal(c.dt, [1, 1, 0.5])
    assert c.tau == 2.5 and cmap[0] == {'X': 'X'} and nmap[1] == {'Z': 'Z', 'Xn': 'Xn'}
    # same identifier, different operator -> position suffix
    b2 = ff.PulseSequence([[Y/2, [3], 'X']], [[Z/2, [1], 'Z']], [0.5])
    c2, cmap2, _ = concatenate_without_filter_function([a, b2], return_identifier_mappings=True)
    assert sorted(c2.c_oper_identifiers) == ['X_0', 'X_1'] and cmap2[1] == {'X': 'X_1'}
    # same operator, different identifiers -> error
    b3 = ff.PulseSequence([[X/2, [3], 'W']], [[Z/2, [1], 'Z']], [0.5])
    with pytest.raises(ValueError):
        concatenate_without_filter_function([a, b3])
    # non-constant sensitivity cannot be inferred
    a4 = ff.PulseSequence([[X/2, [1, 2], 'X']], [[Z/2, [1, 1], 'Z'], [Y/2, [1, 2], 'Yn']], [1, 1])
    with pytest.raises(ValueError):
        concatenate_without_filter_function([a4, b])
    with pytest.raises(TypeError):
        concatenate_without_filter_function([a, 1])
    with pytest.raises(TypeError):
        concatenate_without_filter_function(1)
    d3 = ff.PulseSequence([[np.eye(3), [1]]], [[np.eye(3), [1]]], [1])
    with pytest.raises(ValueError):
        concatenate_without_filter_function([a, d3])
    pb = ff.PulseSequence([[X/2, [1], 'X']], [[Z/2, [1], 'Z']], [1],
                          ff.Basis(np.asarray(ff.Basis.pauli(1))[[0, 2, 1, 3]]))
    with pytest.raises(ValueError):
        concatenate_without_filter_function([a, pb])
    # no cached control matrices and nothing forced: concatenate returns the bare pulse
    plain = ff.concatenate([a, b])
    assert not plain.is_cached('control_matrix')
    with pytest.raises(ValueError):
        ff.concatenate([a, b], calc_filter_function=True)
    with pytest.raises(ValueError):
        ff.concatenate([a, b], which='foo')


def test_basis_constructors_match_oracle_and_properties():
    for n in (1, 2):
        b = ff.Basis.pauli(n)
        np.testing.assert_allclose(np.asarray(b), oracle.pauli_basis(n), atol=1e-15)
        assert b.isherm and b.isorthonorm and b.istraceless and b.iscomplete and b.btype == 'Pauli'
    for d in (2, 3, 5):
        b = ff.Basis.ggm(d)
        np.testing.assert_allclose(np.asarray(b), oracle.ggm_basis(d), atol=1e-15)
        assert b.isherm and b.isorthonorm and b.istraceless and b.iscomplete and b.btype == 'GGM'
        M = np.random.default_rng(d).standard_normal((d, d)) + 0j
        np.testing.assert_allclose(np.einsum('k,kab->ab', b.expand(M), np.asarray(b)), M, atol=1e-14)
    custom = ff.Basis([np.eye(2), np.array([[0, 1], [0, 0]])])
    assert not custom.isherm and not custom.iscomplete and custom.btype == 'Custom'
    with pytest.raises(ValueError):
        ff.Basis(np.zeros((5, 2, 2)))
    with pytest.raises(TypeError):
        ff.Basis(1)
    assert ff.Basis.pauli(1).four_element_traces.shape == (4, 4, 4, 4)


def test_util_parsers():
    omega = np.linspace(1, 2, 5)
    assert util.parse_spectrum(np.ones(5), omega, [0, 1]).shape == (5,)
    assert util.parse_spectrum(np.ones((2, 5)), omega, [0, 1]).shape == (2, 5)
    with pytest.raises(ValueError):
        util.parse_spectrum(np.ones((3, 5)), omega, [0, 1])
    with pytest.raises(ValueError):
        util.parse_spectrum(np.ones(4), omega, [0])
    S = np.ones((2, 2, 5), dtype=complex)
    S[0, 1] += 1j
    with pytest.raises(ValueError):
        util.parse_spectrum(S, omega, [0, 1])
    with pytest.raises(ValueError):
        util.parse_spectrum(np.ones((1, 1, 1, 5)), omega, [0])
    np.testing.assert_array_equal(util.get_indices_from_identifiers(['a', 'b', 'c'], ['c', 'a']),
                                  [2, 0])
    np.testing.assert_array_equal(util.get_indices_from_identifiers(['a', 'b'], 'b'), [1])
    np.testing.assert_array_equal(util.get_indices_from_identifiers(['a', 'b'], None), [0, 1])
    with pytest.raises(ValueError):
        util.get_indices_from_identifiers(['a'], ['z'])
    p = simple_pulse()
    np.testing.assert_allclose(util.get_sample_frequencies(p),
                               oracle.sample_frequencies(p.tau, p.dt.min()))
    w = util.get_sample_frequencies(p, 10, 'linear', include_quasistatic=True)
    assert w[0] == 0 and len(w) == 10
    with pytest.raises(ValueError):
        util.get_sample_frequencies(p, spacing='cubic')
    mats = np.random.default_rng(0).standard_normal((4, 3, 3))
    np.testing.assert_allclose(util.mdot(mats), mats[0] @ mats[1] @ mats[2] @ mats[3])
    np.testing.assert_allclose(util.adot(mats)[-1], mats[3] @ mats[2] @ mats[1] @ mats[0])
    f = np.random.default_rng(1).standard_normal((2, 9))
    x = np.sort(np.random.default_rng(2).random(9))
    np.testing.assert_allclose(util.integrate(f, x), oracle.integrate(f, x))


def test_infidelity_argument_errors_without_gpu():
    p = simple_pulse()
    with pytest.raises(TypeError):
        ff.infidelity(p, 2, [1.0], test_convergence=True)
    with pytest.raises(TypeError):
        ff.infidelity(p, lambda x: x, 2, test_convergence=True)
    with pytest.raises(ValueError):
        ff.infidelity(p, lambda x: x, {'spacing': 'cubic'}, test_convergence=True)
    with pytest.raises(ValueError):
        ff.infidelity(p, [1.0], [1.0], which='foo')
    with pytest.raises(ValueError):
        ff.infidelity(p, [1.0], [1.0], n_oper_identifiers=['nope'])


@pytest.mark.parametrize('n_omega,world', [(1, 1), (1, 4), (2, 2), (5, 8), (10, 3), (10001, 8),
                                           (300, 7), (64, 64)])
def test_frequency_shard_partition(n_omega, world):
    """Intervals are partitioned exactly once; every shard has its halo point."""
    covered = np.zeros(max(n_omega - 1, 0), dtype=int)
    owned = np.zeros(n_omega, dtype=int)
    for r in range(world):
        a, b = ffd.frequency_shard(n_omega, r, world)
        assert 0 <= a <= b <= n_omega
        if b - a >= 2:
            covered[a:b - 1] += 1
        o0, o1 = ffd.owned_frequencies(n_omega, r, world)
        owned[o0:o1] += 1
    assert (covered == 1).all()
    assert (owned == 1).all()
    with pytest.raises(ValueError):
        ffd.frequency_shard(10, 3, 3)


def test_concatenate_hamiltonians_randomized():
    """Long random sequences of recurring pulse objects (the randomized-benchmarking pattern the join
    is written for): every pulse's coefficients must sit in its time slice of the row its (possibly
    renamed) identifier maps to, missing control operators are zero there, missing noise operators
    carry the inferred constant."""
    rng = np.random.default_rng(42)

    def mk(c_ops, n_ops, G):
        return ff.PulseSequence([[o, rng.standard_normal(G), i] for o, i in c_ops],
                                [[o, np.full(G, c), i] for o, i, c in n_ops], np.full(G, 0.3))
    lib = [mk([(X/2, 'X')], [(Z/2, 'Z', 1.0)], 1), mk([(Y/2, 'Y')], [(Z/2, 'Z', 1.0)], 2),
           mk([(X/2, 'X'), (Y/2, 'Y')], [(Z/2, 'Z', 1.0), (X/2, 'Xn', 2.0)], 3),
           mk([(Z/2, 'X')], [(Z/2, 'Z', 1.0)], 1),      # 'X' names another operator here
           mk([(Y/2, 'Y'), (X/2, 'X')], [(X/2, 'Xn', 2.0), (Z/2, 'Z', 1.0)], 2)]
    for _ in range(60):
        seq = [lib[i] for i in rng.integers(0, len(lib), rng.integers(1, 40))]
        new, cmap, nmap = concatenate_without_filter_function(seq, return_identifier_mappings=True)
        edges = np.concatenate(([0], np.cumsum([len(p.dt) for p in seq])))
        assert len(new.dt) == edges[-1]
        for kind, ids, coeffs, maps in (('c', new.c_oper_identifiers, new.c_coeffs, cmap),
                                        ('n', new.n_oper_identifiers, new.n_coeffs, nmap)):
            ids = list(ids)
            assert ids == sorted(ids) and len(set(ids)) == len(ids)
            new_opers = getattr(new, f'{kind}_opers')
            for pos, p in enumerate(seq):
                lo, hi = edges[pos], edges[pos + 1]
                carried = set()
                for ident, op, c in zip(getattr(p, f'{kind}_oper_identifiers'),
                                        getattr(p, f'{kind}_opers'), getattr(p, f'{kind}_coeffs')):
                    # (rows are found by operator: like the reference, the identifier mapping is only
                    # rewritten for the pulse in which a clashing identifier occurs first)
                    row = [k for k in range(len(ids)) if np.array_equal(new_opers[k], op)]
                    assert len(row) == 1
                    carried.add(row[0])
                    np.testing.assert_array_equal(coeffs[row[0], lo:hi], c)
                    assert ids[row[0]] == ident or ids[row[0]].startswith(f'{ident}_')
                    assert maps[pos][ident] in (ident, ids[row[0]])
                for row in set(range(len(ids))) - carried:
                    want = 0.0 if kind == 'c' else coeffs[row][0]
                    assert (coeffs[row, lo:hi] == want).all()


def test_util_tensor_and_float_cleanup():
    """util.tensor: Kronecker product over the last `rank` axes with broadcasting over the leading ones
    (shape rules of the reference's docstring, util.py:367-378); util.remove_float_errors."""
    rng = np.random.default_rng(5)
    Z = np.diag([1, -1])
    assert np.array_equal(util.tensor(Z, Z), np.kron(Z, Z))
    A, B, C = (rng.standard_normal((2, 2)) + 1j*rng.standard_normal((2, 2)) for _ in range(3))
    assert np.allclose(util.tensor(A, B, C), np.kron(np.kron(A, B), C))
    assert np.array_equal(util.tensor(np.arange(2), np.arange(2, 5), rank=1), [0, 0, 0, 2, 3, 4])
    shapes = [(((3, 4, 5, 2, 2), (3, 4, 5, 3, 3), 2), (3, 4, 5, 6, 6)),
              (((7, 2, 3), (7, 4, 5), 2), (7, 8, 15)),
              (((2, 3), (6, 4, 5), 2), (6, 8, 15)),
              (((1, 3), (6, 1, 4), 2), (6, 1, 12)),
              (((5, 2), (5, 3), 1), (5, 6)),
              (((3, 1, 2), (2, 2, 2), 3), (6, 2, 4))]
    for (sa, sb, rank), want in shapes:
        a, b = rng.standard_normal(sa), rng.standard_normal(sb)
        out = util.tensor(a, b, rank=rank)
        assert out.shape == want
    stack = rng.standard_normal((4, 10, 3, 2))
    assert util.tensor(*stack, rank=1).shape == (10, 3, 16)
    assert util.tensor(*stack, rank=2).shape == (10, 81, 16)
    batch = util.tensor(stack[0], stack[1])
    for i in range(10):
        assert np.allclose(batch[i], np.kron(stack[0][i], stack[1][i]))
    with pytest.raises(ValueError, match='Incompatible shapes'):
        util.tensor(rng.standard_normal((3, 1, 2)), rng.standard_normal((2, 2, 2)))
    with pytest.raises(TypeError):
        util.tensor()
    x = np.eye(3) + 1e-17*rng.standard_normal((3, 3))
    assert np.array_equal(util.remove_float_errors(x), np.eye(3))
    y = (1 + 1e-18j)*np.eye(2, dtype=complex)
    assert np.array_equal(util.remove_float_errors(y), np.eye(2))
    assert list(util.progressbar(range(3))) == [0, 1, 2]


def test_basis_from_partial():
    """Basis.from_partial (reference basis.py:492-620): given elements kept (normalised, in order), the
    result complete, orthonormal and Hermitian, identity first for traceless bases, labels, errors."""
    rng = np.random.default_rng(11)
    for d in (2, 3, 4):
        ggm = np.asarray(ff.Basis.ggm(d))
        for n in (1, d, d*d - 1):
            q, _ = np.linalg.qr(rng.standard_normal((d*d - 1, d*d - 1)))
            part = 1.7*np.einsum('ij,jkl->ikl', q[:n], ggm[1:])          # traceless, orthogonal, norm 1.7
            full = ff.Basis.from_partial(part)
            assert full.shape == (d*d, d, d) and full.btype == 'From partial'
            assert full.isherm and full.isorthonorm and full.iscomplete and full.istraceless
            assert np.allclose(np.asarray(full)[0], np.eye(d)/np.sqrt(d))
            assert np.allclose(np.asarray(full)[1:n + 1], part/1.7, atol=1e-13)
        # not traceless: elements stay in front, no identity is inserted
        part = np.zeros((2, d, d))
        part[0, 0, 0] = part[1, 1, 1] = 1
        full = ff.Basis.from_partial(part)
        assert not full.istraceless and full.isorthonorm and full.iscomplete
        assert np.allclose(np.asarray(full)[:2], part)
        with pytest.raises(ValueError, match='traceless'):
            ff.Basis.from_partial(part, traceless=True)
    X, Y, Z = util.paulis[1:]
    named = ff.Basis.from_partial([X, Y], labels=['X', 'Y'])
    assert list(named.labels) == ['X', 'Y', '$C_{2}$', '$C_{3}$']     # as the reference labels them
    with pytest.raises(ValueError, match='not orthogonal'):
        ff.Basis.from_partial([X, X + Y])
    with pytest.raises(ValueError, match='labels'):
        ff.Basis.from_partial([X, Y], labels=['a'])
    assert np.array_equal(ff.Basis.pauli(1).four_element_traces.todense(),
                          np.asarray(ff.Basis.pauli(1).four_element_traces))


def test_choi_and_complete_positivity():
    """superoperator.liouville_to_choi / liouville_is_CP / liouville_is_cCP (reference
    superoperator.py:87-256) on channels with known answers: a unitary channel is CP with Choi
    eigenvalues (d, 0, ...), minus the identity map is not; a Lindblad-type generator is cCP but not CP."""
    from filter_functions_b200 import superoperator as so
    rng = np.random.default_rng(2)
    for d in (2, 3):
        basis = ff.Basis.ggm(d)
        C = np.asarray(basis)
        H = rng.standard_normal((d, d)) + 1j*rng.standard_normal((d, d))
        w, v = np.linalg.eigh(H + H.conj().T)
        U = (v*np.exp(-1j*w)) @ v.conj().T
        L = np.einsum('iab,bc,jcd,da->ij', C, U, C, U.conj().T).real     # tr(C_i U C_j U^dagger)
        choi = so.liouville_to_choi(L, basis)
        direct = sum(L[i, j]*np.kron(C[j].T, C[i]) for i in range(d*d) for j in range(d*d))
        assert np.allclose(choi, direct)
        ok, (vals, _) = so.liouville_is_CP(L, basis, return_eig=True)
        assert ok and np.allclose(sorted(vals)[-1], d) and np.allclose(sorted(vals)[:-1], 0, atol=1e-12)
        assert not so.liouville_is_CP(-np.eye(d*d), basis)
        # dephasing generator  K = sum_k (L_k (x) L_k^* - ...) in Liouville form:  D[rho] = Z rho Z - rho
        Zd = np.diag(np.arange(d) - (d - 1)/2)
        gen = np.einsum('iab,bc,jcd,da->ij', C, Zd, C, Zd).real - 0.5*np.einsum(
            'iab,jbc,ca->ij', C, C, Zd @ Zd).real - 0.5*np.einsum('iab,bc,jca->ij', C, Zd @ Zd, C).real
        assert so.liouville_is_cCP(gen, basis) and not so.liouville_is_CP(gen, basis)
        stack = np.stack([L, -np.eye(d*d)])
        assert list(so.liouville_is_CP(stack, basis)) == [True, False]


def test_join_plan_of_a_superset_is_only_borrowed_when_valid():
    """A sequence over a SUBSET of a gate library uses the remembered plan of the whole library if (and only
    if) its gates together still carry all operators of that plan's join; the result must be what a join from
    scratch gives (same operators, identifiers, coefficient rows, mappings, tau)."""
    from filter_functions_b200 import pulse_sequence as ps
    rng = np.random.default_rng(17)
    X, Y, Z = util.paulis[1:]
    ops = {'X': X/2, 'Y': Y/2, 'Z': Z/2}

    def gate(c_names, n_names, G):
        return ff.PulseSequence([[ops[k], rng.standard_normal(G), k] for k in c_names],
                                [[ops[k], np.ones(G), 'n' + k] for k in n_names], 1 - 0.5*rng.random(G))
    library = [gate('XY', 'Z', 2), gate('X', 'Z', 1), gate('Y', 'ZX', 3), gate('XY', 'ZX', 2), gate('Y', 'Z', 1)]

    def joined(seq):
        new, cmap, nmap = concatenate_without_filter_function(seq, return_identifier_mappings=True)
        return (new.c_opers, list(new.c_oper_identifiers), new.c_coeffs, new.n_opers,
                list(new.n_oper_identifiers), new.n_coeffs, new.dt, new.tau, cmap, nmap)

    ps._SEQUENCE_PLANS.clear()
    joined(library)                                   # remembers the plan of the whole library
    assert len(ps._SEQUENCE_PLANS) == 1
    full = dict(ps._SEQUENCE_PLANS)
    n_borrowed = 0
    for trial in range(60):
        picks = rng.integers(0, len(library), size=rng.integers(2, 9))
        seq = [library[i] for i in picks]
        ps._SEQUENCE_PLANS.clear()
        ps._SEQUENCE_PLANS.update(full)               # only the whole library's plan is known
        distinct, _ = ps._distinct_pulses(seq)
        plan = ps._plan_for(distinct)[1]
        carries_all = ({k for p in distinct for k in p.c_oper_identifiers} == {'X', 'Y'}
                       and {k for p in distinct for k in p.n_oper_identifiers} == {'nX', 'nZ'})
        assert (plan is not None) == carries_all
        n_borrowed += plan is not None
        got = joined(seq)
        ps._SEQUENCE_PLANS.clear()
        want = joined(seq)                            # from scratch
        for a, b in zip(got, want):
            if isinstance(a, np.ndarray):
                assert a.shape == b.shape and np.array_equal(a, b)
            else:
                assert a == b
    assert 5 < n_borrowed < 60
    ps._SEQUENCE_PLANS.clear()
