"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container, where /root/reference exists (it is not shipped to the GPU box).
The reference needs ``opt_einsum`` and ``sparse``; ``oracle/shim`` holds NumPy-backed stand-ins
(SURVEY.md section 8c).  Usage::

    python oracle/gen_golden.py

Every fixture stores the INPUTS next to the reference's OUTPUTS so that the tests need neither the
reference nor its random-number stream.  All results come from the reference's public API
(``ff.PulseSequence``, ``ff.infidelity``, ``ff.concatenate``, ``numeric.*``).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, 'shim'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, '/root/reference/tests')

import filter_functions as ff  # noqa: E402
from filter_functions import numeric, util  # noqa: E402
from tests import testutil  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def pulse_arrays(pulse):
    return dict(c_opers=pulse.c_opers, c_ids=np.asarray(pulse.c_oper_identifiers, dtype='U8'),
                c_coeffs=pulse.c_coeffs, n_opers=pulse.n_opers,
                n_ids=np.asarray(pulse.n_oper_identifiers, dtype='U8'), n_coeffs=pulse.n_coeffs,
                dt=pulse.dt, basis=np.asarray(pulse.basis))


def seeded_infidelities():
    """The seeded regression of tests/test_precision.py:495-551 (seed 123456789)."""
    rng = np.random.default_rng(seed=123456789)
    spectra = [
        lambda S0, omega: S0*abs(omega)**0,
        lambda S0, omega: S0/abs(omega)**0.7,
        lambda S0, omega: S0*np.exp(-abs(omega)),
        lambda S0, omega: np.array([S0*abs(omega)**0, S0/abs(omega)**0.7]),
        lambda S0, omega: np.array([[S0/abs(omega)**0.7, S0/(1 + omega**2) + 1j*S0*omega],
                                    [S0/(1 + omega**2) - 1j*S0*omega, S0/abs(omega)**0.7]])
    ]
    out = {}
    for d in (2, 3, 4):
        pulse = testutil.rand_pulse_sequence(d, 10, 2, 3, local_rng=rng)
        pulse.n_oper_identifiers = np.array(['B_0', 'B_2'])
        omega = np.geomspace(0.1, 10, 51)
        S0 = np.abs(rng.standard_normal())
        for key, val in pulse_arrays(pulse).items():
            if key != 'n_ids':
                out[f'd{d}_{key}'] = val
        out[f'd{d}_omega'] = omega
        out[f'd{d}_S0'] = S0
        for i, spec in enumerate(spectra):
            S = spec(S0, omega)
            out[f'd{d}_spectrum{i}'] = S
            out[f'd{d}_infid{i}'] = ff.infidelity(pulse, S, omega,
                                                   n_oper_identifiers=['B_0', 'B_2'])
        out[f'd{d}_control_matrix'] = pulse.get_control_matrix(omega)
        out[f'd{d}_filter_function'] = pulse.get_filter_function(omega)
    np.savez_compressed(os.path.join(OUT, 'infidelity_seeded.npz'), **out)


def random_pulses():
    """Path results for random pulses of several shapes, incl. non-integer-power-of-two d."""
    rng = np.random.default_rng(20261017)
    out = {}
    cases = [(2, 12, 3, 'Pauli'), (3, 9, 2, 'GGM'), (4, 14, 4, 'Pauli'), (5, 6, 2, 'GGM')]
    for d, G, n_nops, btype in cases:
        pulse = testutil.rand_pulse_sequence(d, G, 3, n_nops, btype=btype, local_rng=rng)
        omega = np.concatenate(([0.0], np.geomspace(1e-3, 40, 60), -np.geomspace(1e-2, 5, 4)))
        tag = f'd{d}'
        for key, val in pulse_arrays(pulse).items():
            out[f'{tag}_{key}'] = val
        out[f'{tag}_omega'] = omega
        pulse.diagonalize()
        out[f'{tag}_eigvals'] = pulse.eigvals
        out[f'{tag}_propagators'] = pulse.propagators
        out[f'{tag}_control_matrix'] = pulse.get_control_matrix(omega)
        out[f'{tag}_filter_function'] = pulse.get_filter_function(omega)
        if d <= 3:  # (n_nops, n_nops, d^2, d^2, n_omega): keep the fixture small
            out[f'{tag}_filter_function_gen'] = pulse.get_filter_function(omega, which='generalized')
        out[f'{tag}_total_phases'] = pulse.get_total_phases(omega)
        out[f'{tag}_total_propagator_liouville'] = pulse.total_propagator_liouville
        om = np.geomspace(0.05, 20, 80)
        S = 1e-3/om**0.7
        out[f'{tag}_int_omega'] = om
        out[f'{tag}_int_spectrum'] = S
        out[f'{tag}_infidelity'] = ff.infidelity(pulse, S, om)
    np.savez_compressed(os.path.join(OUT, 'random_pulses.npz'), **out)


def concatenation():
    """concatenate() of pulses with cached control matrices, equal and differing noise operators,
    pulse-correlation filter functions (cf. tests/test_sequencing.py:222-606, :690-799)."""
    rng = np.random.default_rng(77)
    out = {}
    d = 2
    omega = np.geomspace(1e-2, 30, 70)
    X, Y, Z = util.paulis[1:]
    pulses = []
    for i in range(4):
        G = int(rng.integers(2, 7))
        H_c = [[X/2, rng.standard_normal(G), 'X'], [Y/2, rng.standard_normal(G), 'Y']]
        H_n = [[Z/2, np.ones(G), 'Z'], [X/2, np.full(G, 0.5), 'Xn']]
        if i == 2:  # one pulse lacks a noise operator with constant sensitivity elsewhere
            H_n = H_n[:1]
        pulses.append(ff.PulseSequence(H_c, H_n, 1 - rng.random(G), ff.Basis.pauli(1)))
    for i, pls in enumerate(pulses):
        for key, val in pulse_arrays(pls).items():
            out[f'p{i}_{key}'] = val
        pls.cache_filter_function(omega)
    out['omega'] = omega
    total = ff.concatenate(pulses, calc_pulse_correlation_FF=True)
    out['n_ids'] = np.asarray(total.n_oper_identifiers, dtype='U8')
    out['control_matrix'] = total.get_control_matrix(omega)
    out['control_matrix_pc'] = total.get_pulse_correlation_control_matrix()
    out['filter_function'] = total.get_filter_function(omega)
    out['filter_function_pc'] = total.get_pulse_correlation_filter_function()
    out['total_propagator'] = total.total_propagator
    S = 1e-2/omega
    out['infidelity'] = ff.infidelity(total, S, omega)
    out['infidelity_pc'] = ff.infidelity(total, S, omega, which='correlations')
    # from scratch on the concatenated pulse
    scratch = ff.concatenate(pulses, calc_filter_function=False)
    out['control_matrix_scratch'] = scratch.get_control_matrix(omega)
    np.savez_compressed(os.path.join(OUT, 'concatenation.npz'), **out)


def workloads_small():
    """Reduced-size instances of BASELINE.json configs 2 and 3 built by the repo's own generators, with
    the reference's results (so the generators and the engine are pinned on the bench workloads)."""
    sys.path.insert(0, ROOT)
    import workloads
    out = {}
    for name, kwargs in (('c2', dict(G=64, n_omega=96)), ('c3', dict(G=40, n_omega=64))):
        wl = workloads.get(name, **kwargs)
        pulse = ff.PulseSequence(
            [[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
            [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
            wl.dt, ff.Basis.pauli(int(np.log2(wl.d))))
        out[f'{name}_n_ids'] = np.asarray(pulse.n_oper_identifiers, dtype='U8')
        out[f'{name}_control_matrix'] = pulse.get_control_matrix(wl.omega)
        out[f'{name}_filter_function'] = pulse.get_filter_function(wl.omega)
        out[f'{name}_infidelity'] = ff.infidelity(pulse, wl.spectrum, wl.omega)
    wl = workloads.get('c1')
    pulse = ff.PulseSequence(
        [[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
        [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)], wl.dt)
    out['c1_infidelity'] = ff.infidelity(pulse, wl.spectrum, wl.omega)
    out['c1_filter_function'] = pulse.get_filter_function(wl.omega)
    np.savez_compressed(os.path.join(OUT, 'workloads_small.npz'), **out)


def resonance_frequency_picks(wl, n_pick=1024, seed=5):
    """Indices into ``wl.omega`` for the full-size parity fixtures: both ends of the grid, the
    neighbourhoods of the level splittings |E_m - E_n| of a sample of segments (where the first-order
    integral goes through its removable singularity and the kernels switch to their fix-up branch),
    and a uniform random remainder.  Deterministic given the workload."""
    rng = np.random.default_rng(seed)
    n = len(wl.omega)
    H = np.einsum('ijk,il->ljk', wl.c_opers, wl.c_coeffs[:, ::max(1, wl.G//64)])
    ev = np.linalg.eigvalsh(H)
    gaps = np.abs(ev[:, :, None] - ev[:, None, :])
    gaps = np.unique(gaps[gaps > 0])
    near = np.searchsorted(wl.omega, gaps).clip(1, n - 2)
    picks = [np.arange(48), np.arange(n - 48, n)]
    picks += [near + k for k in (-1, 0, 1)]
    have = np.unique(np.concatenate(picks).clip(0, n - 1))
    if len(have) > n_pick - 64:
        have = np.sort(rng.choice(have, n_pick - 64, replace=False))
    rest = np.setdiff1d(np.arange(n), have)
    extra = rng.choice(rest, n_pick - len(have), replace=False)
    return np.sort(np.concatenate([have, extra]))


def workloads_full():
    """The bench workloads at BASELINE.json's FULL sizes, computed once by the unmodified reference
    (config 2: ~70 s, config 3 and the north_star d = 4 shape: a few minutes each on 8 cores).  Config 2
    keeps its whole control matrix (1.9 MB); the d = 4 shapes keep the control matrix on 1024 selected
    frequencies (both grid ends, resonance neighbourhoods, random rest) and the infidelities of the
    whole grid.  The inputs come from workloads.py (seeded); a checksum of them is stored."""
    import hashlib
    import time
    sys.path.insert(0, ROOT)
    import workloads
    names = sys.argv[2:] or ['c2', 'c3', 'd4']
    for name in names:
        wl = workloads.get(name)
        pulse = ff.PulseSequence(
            [[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
            [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
            wl.dt, ff.Basis.pauli(int(np.log2(wl.d))))
        t0 = time.perf_counter()
        B = pulse.get_control_matrix(wl.omega)
        seconds = time.perf_counter() - t0
        digest = hashlib.sha256(np.ascontiguousarray(wl.c_coeffs).tobytes()
                                + np.ascontiguousarray(wl.n_coeffs).tobytes()
                                + np.ascontiguousarray(wl.omega).tobytes()).hexdigest()
        out = dict(n_ids=np.asarray(pulse.n_oper_identifiers, dtype='U8'),
                   infidelity=ff.infidelity(pulse, wl.spectrum, wl.omega),
                   scale=np.abs(B).max(axis=(1, 2)), input_sha256=np.asarray(digest),
                   reference_seconds=seconds, total_propagator=pulse.total_propagator)
        if name == 'c2':
            out['pick'] = np.arange(len(wl.omega))
            out['control_matrix'] = B
        else:
            pick = resonance_frequency_picks(wl)
            out['pick'] = pick
            out['control_matrix'] = np.ascontiguousarray(B[..., pick])
        F = pulse.get_filter_function(wl.omega)
        out['filter_function'] = np.ascontiguousarray(F[..., out['pick']])
        np.savez_compressed(os.path.join(OUT, f'workload_full_{name}.npz'), **out)
        print(name, f'{seconds:.1f} s reference control matrix', flush=True)


def decay_amplitudes():
    """Decay amplitudes, cumulant function and error transfer matrix (numeric.py:957-1337, :1938-2059;
    cf. tests/test_precision.py:631-727) for the three spectrum shapes, plus the pulse-correlation
    amplitudes of a concatenated sequence."""
    rng = np.random.default_rng(314159)
    out = {}
    for d, G, n_nops, btype in [(2, 8, 3, 'Pauli'), (3, 7, 2, 'GGM'), (4, 9, 3, 'Pauli')]:
        pulse = testutil.rand_pulse_sequence(d, G, 2, n_nops, btype=btype, local_rng=rng)
        omega = np.geomspace(0.05, 30, 73)
        tag = f'd{d}'
        for key, val in pulse_arrays(pulse).items():
            out[f'{tag}_{key}'] = val
        out[f'{tag}_omega'] = omega
        S1 = 1e-2/omega**0.7
        S2 = np.array([S1*(i + 1) for i in range(n_nops)])
        S3 = np.einsum('a,b,o->abo', np.arange(1, n_nops + 1), np.arange(1, n_nops + 1), S1)
        S3 = S3 + 1j*np.triu(np.ones((n_nops, n_nops)), 1)[..., None]*omega*1e-3 \
            - 1j*np.tril(np.ones((n_nops, n_nops)), -1)[..., None]*omega*1e-3
        for i, S in enumerate((S1, S2, S3)):
            out[f'{tag}_spectrum{i}'] = S
            out[f'{tag}_decay_amplitudes{i}'] = numeric.calculate_decay_amplitudes(pulse, S, omega)
            out[f'{tag}_cumulant_function{i}'] = numeric.calculate_cumulant_function(pulse, S, omega)
            out[f'{tag}_error_transfer_matrix{i}'] = ff.error_transfer_matrix(pulse, S, omega)
        # subset of noise operators
        ids = pulse.n_oper_identifiers[[0, n_nops - 1]]
        out[f'{tag}_subset_ids'] = np.asarray(ids, dtype='U8')
        out[f'{tag}_decay_amplitudes_subset'] = numeric.calculate_decay_amplitudes(
            pulse, S2[[0, n_nops - 1]], omega, n_oper_identifiers=ids)
        # pulse correlations: the same pulse split in two and concatenated
        halves = [pulse[:G//2], pulse[G//2:]]
        for h in halves:
            h.cache_control_matrix(omega)
        seq = ff.concatenate(halves, calc_pulse_correlation_FF=True)
        out[f'{tag}_split'] = np.array(G//2)
        for i, S in enumerate((S1, S2, S3)):
            out[f'{tag}_decay_amplitudes_pc{i}'] = numeric.calculate_decay_amplitudes(
                seq, S, omega, which='correlations')
        out[f'{tag}_cumulant_function_pc0'] = numeric.calculate_cumulant_function(
            seq, S1, omega, which='correlations')
    np.savez_compressed(os.path.join(OUT, 'decay_amplitudes.npz'), **out)


def sequencing_workloads():
    """BASELINE.json configs 4 and 5 at reduced size, built by workloads.py through the reference's
    public API: randomized-benchmarking sequences of cached Cliffords (examples/
    randomized_benchmarking.py) and the QFT concatenated from cached gate pulses (examples/qft.py)."""
    sys.path.insert(0, ROOT)
    import workloads
    out = {}
    # ---- C4
    omega = workloads.rb_omega()
    S = workloads.rb_spectrum(omega)
    cliffords = workloads.build_cliffords(ff, omega)
    out['rb_omega'] = omega
    out['rb_spectrum'] = S
    out['rb_clifford_infidelity'] = np.array([ff.infidelity(c, S, omega) for c in cliffords])
    out['rb_clifford_propagators'] = np.array([c.total_propagator for c in cliffords])
    rows = workloads.rb_sequences(6, 100)
    out['rb_rows'] = rows
    seqs = [ff.concatenate([cliffords[k] for k in row]) for row in rows]
    out['rb_infidelity'] = np.array([ff.infidelity(p, S, omega) for p in seqs])
    out['rb_filter_function'] = np.array([p.get_filter_function(omega) for p in seqs])
    out['rb_control_matrix'] = np.array([p.get_control_matrix(omega) for p in seqs])
    out['rb_total_propagator'] = np.array([p.total_propagator for p in seqs])
    # ---- C5
    for N, n_omega in ((2, 60), (3, 40), (4, 24)):
        omega = np.logspace(-2, 2, n_omega)
        pulses = workloads.build_qft_pulses(ff, N)
        for p in pulses:
            p.cache_control_matrix(omega)
        qft = ff.concatenate(pulses, omega=omega)
        B = qft.get_control_matrix(omega)
        tag = f'qft{N}'
        out[f'{tag}_omega'] = omega
        out[f'{tag}_n_ids'] = np.asarray(qft.n_oper_identifiers, dtype='U8')
        out[f'{tag}_total_propagator'] = qft.total_propagator
        out[f'{tag}_filter_function'] = qft.get_filter_function(omega)
        out[f'{tag}_control_matrix'] = B if N < 4 else B[:, ::16]
        out[f'{tag}_infidelity'] = ff.infidelity(qft, 1e-4/omega, omega)
        scratch = ff.concatenate(pulses, calc_filter_function=False)
        out[f'{tag}_filter_function_scratch'] = scratch.get_filter_function(omega)
    np.savez_compressed(os.path.join(OUT, 'sequencing_workloads.npz'), **out)


def cnot_gate():
    """The reference's own experimental fixture: the exchange-coupled singlet-triplet CNOT of
    examples/data/CNOT.mat (250 segments of 0.2 ns) on the 6-dimensional subspace with the padded
    two-qubit Pauli basis (15 elements, not complete), as in tests/test_precision.py:184-216 and
    :274-311 (``cnot.d = 4``; Monte-Carlo infidelities testutil.cnot_infid_fast)."""
    c_opers = testutil.subspace_opers
    identifiers = ['eps_12', 'eps_23', 'eps_34', 'b_12', 'b_23', 'b_34']
    basis = ff.Basis([np.pad(b, 1, 'constant') for b in ff.Basis.pauli(2)[1:]], btype='Pauli')
    cnot = ff.PulseSequence(list(zip(c_opers, testutil.c_coeffs, identifiers)),
                            list(zip(c_opers, testutil.n_coeffs, identifiers)), testutil.dt,
                            basis=basis)
    cnot.d = 4
    omega = np.geomspace(1/cnot.tau, 1e2, 250)
    out = {f'cnot_{k}': v for k, v in pulse_arrays(cnot).items()}
    out['cnot_omega'] = omega
    out['cnot_A'] = np.asarray(testutil.A)
    out['cnot_infid_MC'] = np.asarray(testutil.cnot_infid_fast)
    cnot.diagonalize()
    out['cnot_total_propagator'] = cnot.total_propagator
    out['cnot_eigvals'] = cnot.eigvals
    out['cnot_control_matrix'] = cnot.get_control_matrix(omega)
    out['cnot_filter_function'] = cnot.get_filter_function(omega)
    for i, (A, alpha) in enumerate(zip(testutil.A, (0.0, 0.7))):
        infid, xi = ff.infidelity(cnot, A/omega**alpha, omega, identifiers[:3],
                                  return_smallness=True)
        out[f'cnot_infid_{i}'] = infid
        out[f'cnot_xi_{i}'] = xi
    np.savez_compressed(os.path.join(OUT, 'cnot.npz'), **out)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:       # regenerate only the named fixtures
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    seeded_infidelities()
    random_pulses()
    concatenation()
    workloads_small()
    decay_amplitudes()
    sequencing_workloads()
    cnot_gate()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), 'bytes')
