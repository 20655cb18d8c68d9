"""Differential fuzz of the public API against the UNMODIFIED reference (GPU box).

Random small pulses (d = 2 .. 6, 1 .. 40 segments, Pauli / GGM / custom bases, traceless or not, Hermitian
noise operators, uniform and ragged time grids), random frequency grids (log, linear, with zero and
negative frequencies) and spectra (1-d, 2-d, 3-d complex Hermitian), random identifier subsets.  Every
quantity of the path is computed by this package (GPU) and by the reference staged under baseline/_ref
(CPU, NumPy) from the same inputs and compared at the north-star tolerance: 1e-10, normalised by the
largest magnitude of the reference's array.  Prints one JSON line with the number of cases and
comparisons and every mismatch.

    python tools/fuzz_vs_reference.py [--cases 300] [--seed 1]
"""
import argparse
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim'))
sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
warnings.filterwarnings('ignore')

import filter_functions as ref  # noqa: E402  (the reference)
import filter_functions_b200 as ffb  # noqa: E402

TOL = 1e-10


def herm(rng, d, n, traceless):
    a = rng.standard_normal((n, d, d)) + 1j*rng.standard_normal((n, d, d))
    a = a + a.conj().swapaxes(1, 2)
    if traceless:
        a -= np.einsum('nii->n', a)[:, None, None]*np.eye(d)/d
    return a


def make_pulses(rng, d, G, basis_kind):
    n_cops, n_nops = rng.integers(1, 4), rng.integers(1, 5)
    c_opers = herm(rng, d, n_cops, True)
    n_opers = herm(rng, d, n_nops, bool(rng.integers(2)))
    if rng.integers(5) == 0:    # the reference takes arbitrary noise operators, not only Hermitian ones
        n_opers = n_opers + 1j*herm(rng, d, n_nops, True)
    c_coeffs = rng.standard_normal((n_cops, G))
    n_coeffs = rng.random((n_nops, G)) + 0.3 if rng.integers(2) else np.ones((n_nops, G))
    dt = np.full(G, rng.uniform(0.2, 1.5)) if rng.integers(2) else rng.uniform(0.1, 1.2, G)
    c_ids = [f'C{i}' for i in rng.permutation(n_cops)]
    n_ids = [f'N{i}' for i in rng.permutation(n_nops)]
    out = []
    for pkg in (ref, ffb):
        if basis_kind == 'pauli':
            basis = pkg.Basis.pauli(int(np.log2(d)))
        elif basis_kind == 'ggm':
            basis = pkg.Basis.ggm(d)
        elif basis_kind == 'default':
            basis = None
        else:   # partial basis completed by from_partial (not traceless in general)
            basis = pkg.Basis.from_partial(n_opers[:1]/np.linalg.norm(n_opers[0]), traceless=False)
        out.append(pkg.PulseSequence(list(zip(c_opers, c_coeffs, c_ids)), list(zip(n_opers, n_coeffs, n_ids)),
                                     dt, basis))
    return out[0], out[1], n_nops


def make_omega(rng, pulse):
    kind = rng.integers(4)
    n = int(rng.integers(2, 120))
    tau = pulse.dt.sum()
    if kind == 0:
        return np.geomspace(1e-2/tau, 30/pulse.dt.min(), n)
    if kind == 1:
        return np.linspace(0, 20/tau, n)
    if kind == 2:
        return np.sort(rng.uniform(-10/tau, 10/tau, n))
    return ref.util.get_sample_frequencies(pulse, n_samples=n, spacing='log')


def make_spectrum(rng, omega, n_sel):
    kind = rng.integers(3)
    base = 1e-3/(1 + omega**2)
    if kind == 0:
        return base
    if kind == 1:
        return np.outer(rng.uniform(0.5, 2, n_sel), base)
    S = np.zeros((n_sel, n_sel, len(omega)), dtype=complex)
    for a in range(n_sel):
        S[a, a] = rng.uniform(0.5, 2)*base
        for b in range(a + 1, n_sel):
            S[a, b] = (rng.uniform(-0.3, 0.3) + 1j*rng.uniform(-0.3, 0.3)*np.tanh(omega))*base
            S[b, a] = S[a, b].conj()
    return S


def nerr(x, y):
    x, y = np.asarray(x), np.asarray(y)
    if x.shape != y.shape:
        return np.inf
    scale = np.abs(y).max() if y.size else 0.0
    return float(np.abs(x - y).max()/scale) if scale > 0 else float(np.abs(x - y).max() if x.size else 0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--cases', type=int, default=300)
    ap.add_argument('--seed', type=int, default=1)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    mismatches, n_cmp, worst = [], 0, 0.0

    def check(case, what, got, want, tol=TOL):
        nonlocal n_cmp, worst
        n_cmp += 1
        e = nerr(got, want)
        worst = max(worst, e if np.isfinite(e) else 0.0)
        if not e <= tol:
            mismatches.append({'case': case, 'what': what, 'err': e})

    for case in range(args.cases):
        d = int(rng.choice([2, 2, 3, 4, 4, 5, 6]))
        G = int(rng.integers(1, 41))
        kinds = ['ggm', 'default', 'partial'] + (['pauli'] if d in (2, 4) else [])
        kind = kinds[rng.integers(len(kinds))]
        p_ref, p_new, n_nops = make_pulses(rng, d, G, kind)
        omega = make_omega(rng, p_ref)
        tag = f'{case}: d={d} G={G} basis={kind} n_nops={n_nops} n_omega={len(omega)}'
        check(tag, 'control matrix', p_new.get_control_matrix(omega), p_ref.get_control_matrix(omega))
        check(tag, 'filter function', p_new.get_filter_function(omega), p_ref.get_filter_function(omega))
        check(tag, 'total propagator', p_new.total_propagator, p_ref.total_propagator)
        check(tag, 'eigvals', p_new.eigvals, p_ref.eigvals)
        if d <= 4 and len(omega) <= 60:
            check(tag, 'generalized filter function', p_new.get_filter_function(omega, which='generalized'),
                  p_ref.get_filter_function(omega, which='generalized'))
        ids = list(p_ref.n_oper_identifiers)
        sel = sorted(rng.choice(len(ids), size=rng.integers(1, len(ids) + 1), replace=False))
        sel_ids = [ids[i] for i in sel] if rng.integers(2) else None
        S = make_spectrum(rng, omega, len(ids) if sel_ids is None else len(sel_ids))
        check(tag, 'infidelity', ffb.infidelity(p_new, S, omega, n_oper_identifiers=sel_ids),
              ref.infidelity(p_ref, S, omega, n_oper_identifiers=sel_ids))
        if kind != 'partial' or True:
            try:
                want = ref.numeric.calculate_decay_amplitudes(p_ref, S, omega, sel_ids)
            except Exception as exc:   # noqa: BLE001  (the reference refuses: so must we)
                try:
                    ffb.numeric.calculate_decay_amplitudes(p_new, S, omega, sel_ids)
                    mismatches.append({'case': tag, 'what': 'decay amplitudes: reference raised', 'err': repr(exc)})
                except Exception:   # noqa: BLE001
                    pass
            else:
                check(tag, 'decay amplitudes', ffb.numeric.calculate_decay_amplitudes(p_new, S, omega, sel_ids), want)
        if d <= 4 and kind != 'partial':
            for fn in ('calculate_cumulant_function',):
                try:
                    want = getattr(ref.numeric, fn)(p_ref, S, omega, sel_ids)
                except Exception:   # noqa: BLE001
                    continue
                check(tag, fn, getattr(ffb.numeric, fn)(p_new, S, omega, sel_ids), want)
            try:
                want = ref.error_transfer_matrix(p_ref, S, omega, n_oper_identifiers=sel_ids)
            except Exception:   # noqa: BLE001
                want = None
            if want is not None:
                check(tag, 'error transfer matrix',
                      ffb.error_transfer_matrix(p_new, S, omega, n_oper_identifiers=sel_ids), want)
        if len(omega) <= 40 and G <= 12:
            B_new, i_new = ffb.numeric.calculate_control_matrix_from_scratch(
                p_new.eigvals, p_new.eigvecs, p_new.propagators, omega, p_new.basis, p_new.n_opers,
                p_new.n_coeffs, p_new.dt, cache_intermediates=True)
            B_ref, i_ref = ref.numeric.calculate_control_matrix_from_scratch(
                p_new.eigvals, p_new.eigvecs, p_new.propagators, omega, p_ref.basis, p_ref.n_opers,
                p_ref.n_coeffs, p_ref.dt, cache_intermediates=True)
            check(tag, 'control matrix (same eigensystem)', B_new, B_ref)
            for key in i_ref:
                check(tag, 'intermediates/' + key, i_new[key], i_ref[key])
        # concatenation of the pulse with a second one on the same operators
        if rng.integers(2):
            q_ref, q_new = p_ref[:max(1, G//2)], p_new[:max(1, G//2)]
            for a, b in ((p_ref, q_ref), (p_new, q_new)):
                a.cache_filter_function(omega)
                b.cache_filter_function(omega)
            pc = bool(rng.integers(2))
            c_ref = ref.concatenate([p_ref, q_ref, p_ref], calc_pulse_correlation_FF=pc)
            c_new = ffb.concatenate([p_new, q_new, p_new], calc_pulse_correlation_FF=pc)
            check(tag, 'concatenated filter function', c_new.get_filter_function(omega),
                  c_ref.get_filter_function(omega))
            check(tag, 'concatenated control matrix', c_new.get_control_matrix(omega),
                  c_ref.get_control_matrix(omega))
            if pc:
                check(tag, 'pulse correlation filter function', c_new.get_pulse_correlation_filter_function(),
                      c_ref.get_pulse_correlation_filter_function())
                S1 = make_spectrum(rng, omega, len(ids))
                check(tag, 'infidelity correlations', ffb.infidelity(c_new, S1, omega, which='correlations'),
                      ref.infidelity(c_ref, S1, omega, which='correlations'))
            reps = int(rng.integers(2, 6))
            check(tag, 'periodic filter function',
                  ffb.concatenate_periodic(p_new, reps).get_filter_function(omega),
                  ref.concatenate_periodic(p_ref, reps).get_filter_function(omega), tol=1e-9)
    # concatenation of pulses that carry DIFFERENT (overlapping) sets of noise operators, omega supplied or cached
    for case in range(args.cases//4):
        d = int(rng.choice([2, 3, 4]))
        pool = herm(rng, d, 4, True)
        c_opers = herm(rng, d, 2, True)
        sens = rng.uniform(0.5, 2, 4)
        refs, news = [], []
        for k in range(int(rng.integers(2, 5))):
            G = int(rng.integers(1, 9))
            pick = sorted(rng.choice(4, size=rng.integers(1, 5), replace=False))
            const = bool(rng.integers(2))
            n_coeffs = [np.full(G, sens[i]) if const else sens[i]*(1 + 0.1*rng.standard_normal(G)) for i in pick]
            c_coeffs = rng.standard_normal((2, G))
            dt = rng.uniform(0.2, 1.0, G)
            for pkg, dst in ((ref, refs), (ffb, news)):
                dst.append(pkg.PulseSequence(list(zip(c_opers, c_coeffs, ['X', 'Y'])),
                                             [[pool[i], nc, f'N{i}'] for i, nc in zip(pick, n_coeffs)], dt,
                                             pkg.Basis.ggm(d)))
        omega = np.geomspace(1e-2, 20, int(rng.integers(2, 60)))
        tag = f'mixed {case}: d={d} pulses={len(refs)}'
        cached = bool(rng.integers(2))
        try:
            if cached:
                for pr_, pn_ in zip(refs, news):
                    pr_.cache_filter_function(omega)
                    pn_.cache_filter_function(omega)
                c_ref = ref.concatenate(refs, calc_filter_function=True)
            else:
                c_ref = ref.concatenate(refs, omega=omega, calc_filter_function=True)
        except Exception as exc:   # noqa: BLE001
            try:
                ffb.concatenate(news, omega=None if cached else omega, calc_filter_function=True)
                mismatches.append({'case': tag, 'what': 'reference raised, this package did not', 'err': repr(exc)})
            except type(exc):
                n_cmp += 1
            continue
        c_new = ffb.concatenate(news, omega=None if cached else omega, calc_filter_function=True)
        check(tag, 'identifiers', np.array(list(c_new.n_oper_identifiers) == list(c_ref.n_oper_identifiers), float),
              np.ones(()))
        check(tag, 'n_coeffs', c_new.n_coeffs, c_ref.n_coeffs)
        check(tag, 'filter function', c_new.get_filter_function(omega), c_ref.get_filter_function(omega))
        check(tag, 'control matrix', c_new.get_control_matrix(omega), c_ref.get_control_matrix(omega))
        check(tag, 'total propagator', c_new.total_propagator, c_ref.total_propagator)

    print(json.dumps({'cases': args.cases, 'seed': args.seed, 'comparisons': n_cmp, 'tolerance': TOL,
                      'worst_normalised_deviation': worst, 'mismatches': mismatches[:40],
                      'n_mismatches': len(mismatches)}))


if __name__ == '__main__':
    main()
