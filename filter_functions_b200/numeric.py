"""Numerics of the control-matrix / filter-function / infidelity path, executed on the GPU.

Same public names, signatures, defaults, return shapes and error behaviour as the hot-path functions
of the reference's ``numeric.py``; the bodies call ``libffb200`` (hand-written sm_100a kernels) through
ctypes.  NumPy arrays in, NumPy arrays out.  There is no CPU fallback.

=====================================================  ==========================================
this module                                            reference ``numeric.py``
=====================================================  ==========================================
:func:`diagonalize`                                    ``:1886-1935``
:func:`calculate_control_matrix_from_scratch`          ``:707-881``
:func:`calculate_control_matrix_from_atomic`           ``:621-704``
:func:`calculate_filter_function`                      ``:1413-1467``
:func:`calculate_pulse_correlation_filter_function`    ``:1821-1883``
:func:`infidelity`                                     ``:2062-2334``
=====================================================  ==========================================
"""
from typing import Optional

import numpy as np
from numpy import ndarray

from . import _lib, util

__all__ = ['calculate_control_matrix_from_atomic', 'calculate_control_matrix_from_scratch',
           'calculate_control_matrix_periodic', 'calculate_cumulant_function', 'calculate_decay_amplitudes',
           'calculate_filter_function', 'calculate_pulse_correlation_filter_function',
           'diagonalize', 'error_transfer_matrix', 'infidelity']


def diagonalize(hamiltonian: ndarray, dt):
    """Eigenvalues (n_dt, d), eigenvectors (n_dt, d, d) and cumulative propagators (n_dt+1, d, d) of
    a piecewise-constant Hamiltonian of shape (n_dt, d, d)."""
    H = _lib.as_c128(hamiltonian)
    if H.ndim != 3 or H.shape[-1] != H.shape[-2]:
        raise ValueError(f'Expected hamiltonian of shape (n_dt, d, d), not {H.shape}')
    dt = _lib.as_f64(dt)
    G, d = H.shape[0], H.shape[1]
    if dt.shape != (G,):
        raise ValueError(f'Expected dt of shape ({G},), not {dt.shape}')
    eigvals = _lib.empty((G, d), np.float64)
    eigvecs = _lib.empty((G, d, d))
    propagators = _lib.empty((G + 1, d, d))
    ctx = _lib.context()
    _lib.check(ctx, _lib.lib().ffb_diagonalize(ctx, G, d, 0, _lib.ptr(H), None, _lib.ptr(dt),
                                               _lib.ptr(eigvals), _lib.ptr(eigvecs),
                                               _lib.ptr(propagators)))
    return eigvals, eigvecs, propagators


def _diagonalize_from_coeffs(c_opers, c_coeffs, dt):
    """PulseSequence.diagonalize without materialising H on the host (``pulse_sequence.py:577-586``)."""
    c_opers = _lib.as_c128(c_opers)
    c_coeffs = _lib.as_f64(c_coeffs)
    dt = _lib.as_f64(dt)
    n_cops, d = c_opers.shape[0], c_opers.shape[-1]
    G = dt.shape[0]
    eigvals = _lib.empty((G, d), np.float64)
    eigvecs = _lib.empty((G, d, d))
    propagators = _lib.empty((G + 1, d, d))
    ctx = _lib.context()
    _lib.check(ctx, _lib.lib().ffb_diagonalize(ctx, G, d, n_cops, _lib.ptr(c_opers),
                                               _lib.ptr(c_coeffs), _lib.ptr(dt), _lib.ptr(eigvals),
                                               _lib.ptr(eigvecs), _lib.ptr(propagators)))
    return eigvals, eigvecs, propagators


def calculate_control_matrix_from_scratch(eigvals, eigvecs, propagators, omega, basis, n_opers,
                                          n_coeffs, dt, t=None, show_progressbar: bool = False,
                                          cache_intermediates: bool = False,
                                          out: Optional[ndarray] = None):
    r"""Control matrix :math:`\tilde{\mathcal{B}}_{\alpha k}(\omega)` of shape
    (n_nops, n_basis, n_omega) without knowledge of more atomic pulses."""
    return _control_matrix_from_scratch(eigvals, eigvecs, propagators, omega, basis, n_opers,
                                        n_coeffs, dt, t, cache_intermediates, out, keep=False)


def _control_matrix_from_scratch(eigvals, eigvecs, propagators, omega, basis, n_opers, n_coeffs, dt,
                                 t=None, cache_intermediates: bool = False,
                                 out: Optional[ndarray] = None, keep: bool = False):
    """Body of :func:`calculate_control_matrix_from_scratch`.  ``keep``: the result stays mirrored on the
    device (and becomes read-only on the host) -- used by ``PulseSequence``, whose cache hands the array
    to later device computations."""
    eigvals = _lib.as_f64(eigvals)
    eigvecs = _lib.as_c128(eigvecs)
    propagators = _lib.as_c128(propagators)
    omega = _lib.as_f64(np.asanyarray(omega))
    basis_arr = _lib.as_c128(np.asarray(basis))
    n_opers = _lib.as_c128(n_opers)
    n_coeffs = _lib.as_f64(n_coeffs)
    dt = _lib.as_f64(dt)
    if t is None:
        t = np.concatenate(([0], dt.cumsum()))
    t = _lib.as_f64(t)
    G, d = eigvals.shape
    n_nops, n_basis, n_omega = len(n_opers), len(basis_arr), len(omega)
    if eigvecs.shape != (G, d, d) or propagators.shape != (G + 1, d, d):
        raise ValueError('eigvals, eigvecs and propagators have inconsistent shapes')
    if n_coeffs.shape != (n_nops, G) or dt.shape != (G,) or t.shape != (G + 1,):
        raise ValueError('n_coeffs, dt or t do not match the number of segments')
    result = out
    direct = (out is not None and out.dtype == np.complex128 and out.flags.c_contiguous
              and out.shape == (n_nops, n_basis, n_omega))
    if not direct:
        result = _lib.empty((n_nops, n_basis, n_omega))
    ctx = _lib.context()
    if cache_intermediates:
        # materialising variant (numeric.py:828-879): same result plus the (G, n_omega, ...)-sized
        # by-products the fused kernel never builds
        inter = dict(
            n_opers_transformed=_lib.empty((n_nops, G, d, d)),
            eigvecs_propagated=_lib.empty((G, d, d)),
            basis_transformed=_lib.empty((G, n_basis, d, d)),
            phase_factors=_lib.empty((G, n_omega)),
            first_order_integral=_lib.empty((G, n_omega, d, d)),
            control_matrix_step=_lib.empty((G, n_nops, n_basis, n_omega)),
            control_matrix_step_cumulative=_lib.empty((G - 1, n_nops, n_basis, n_omega)))
        if n_omega:
            _lib.check(ctx, _lib.lib().ffb_control_matrix_intermediates(
                ctx, G, d, n_nops, n_basis, n_omega, _lib.ptr(eigvals), _lib.ptr(eigvecs),
                _lib.ptr(propagators), _lib.ptr(omega), _lib.ptr(basis_arr), _lib.ptr(n_opers),
                _lib.ptr(n_coeffs), _lib.ptr(dt), _lib.ptr(t), _lib.ptr(result),
                *(_lib.ptr(inter[key]) for key in (
                    'n_opers_transformed', 'eigvecs_propagated', 'basis_transformed',
                    'phase_factors', 'first_order_integral', 'control_matrix_step',
                    'control_matrix_step_cumulative'))))
        if out is not None and not direct:
            out[:] = result
            result = out
        return result, inter
    if n_omega:
        if keep and out is None:
            _lib.keep_on_device(ctx)
        _lib.check(ctx, _lib.lib().ffb_control_matrix_from_scratch(
            ctx, G, d, n_nops, n_basis, n_omega, _lib.ptr(eigvals), _lib.ptr(eigvecs),
            _lib.ptr(propagators), _lib.ptr(omega), _lib.ptr(basis_arr), _lib.ptr(n_opers),
            _lib.ptr(n_coeffs), _lib.ptr(dt), _lib.ptr(t), _lib.ptr(result)))
        if keep and out is None:
            _lib.freeze_shadowed(ctx, result)
    if out is not None and not direct:
        out[:] = result
        return out
    return result


@util.parse_optional_parameters(which=('total', 'correlations'))
def calculate_control_matrix_from_atomic(phases, control_matrix_atomic, propagators_liouville,
                                         show_progressbar: bool = False, which: str = 'total'):
    r"""Control matrix of a sequence from those of its constituents,
    :math:`\sum_g e^{i\omega t_{g-1}}\tilde{\mathcal{B}}^{(g)}(\omega)\mathcal{Q}^{(g-1)}`.
    The memory order of the result mirrors that of ``control_matrix_atomic`` (C, F or neither), as
    in the reference (``numeric.py:671-676``)."""
    atomic_in = np.asarray(control_matrix_atomic)
    if atomic_in.ndim != 4:
        raise ValueError('Expected control_matrix_atomic.ndim == 4.')
    c_in, f_in = atomic_in.flags.c_contiguous, atomic_in.flags.f_contiguous
    atomic = _lib.as_c128(atomic_in)
    P, n_nops, n_basis, n_omega = atomic.shape
    # the reference reads phases[g-1] and propagators_liouville[g-1] for g = 1 .. P-1 (numeric.py:691-701):
    # arrays with more leading entries than P - 1 are legal, the rest is never looked at
    phases = np.asarray(phases)
    Q = np.asarray(propagators_liouville)
    if phases.ndim != 2 or phases.shape[0] < P - 1 or phases.shape[1] != n_omega:
        raise ValueError(f'Expected phases of shape ({P - 1}, {n_omega}), not {phases.shape}')
    if Q.ndim != 3 or Q.shape[0] < P - 1 or Q.shape[1:] != (n_basis, n_basis):
        raise ValueError(f'Expected propagators_liouville of shape ({P - 1}, {n_basis}, {n_basis}), '
                         f'not {Q.shape}')
    phases = _lib.as_c128(phases[:P - 1])
    q_complex = np.iscomplexobj(Q)
    Q = _lib.as_c128(Q[:P - 1]) if q_complex else _lib.as_f64(Q[:P - 1])
    corr = which == 'correlations'
    out = _lib.empty((P, n_nops, n_basis, n_omega) if corr else (n_nops, n_basis, n_omega))
    ctx = _lib.context()
    _lib.check(ctx, _lib.lib().ffb_control_matrix_from_atomic(
        ctx, P, n_nops, n_basis, n_omega, _lib.ptr(phases), _lib.ptr(atomic), _lib.ptr(Q),
        int(q_complex), int(corr), _lib.ptr(out)))
    if c_in:
        return out
    if f_in:
        return np.asfortranarray(out)
    # neither: the reference hands back an array with omega on the second-to-last memory axis
    return np.ascontiguousarray(out.swapaxes(-1, -2)).swapaxes(-1, -2)


def calculate_control_matrix_periodic(phases, control_matrix, total_propagator_liouville,
                                      repeats: int, check_invertible: bool = True) -> ndarray:
    r"""Control matrix of a pulse repeated ``repeats`` times,
    :math:`\tilde{\mathcal B}(\omega)\sum_{g<G}(e^{i\omega T}\mathcal Q)^g` (reference
    ``numeric.py:884-954``).  The geometric series is evaluated by binary doubling on the GPU, which
    is valid at every frequency, so ``check_invertible`` (the reference's guard around its per-
    frequency linear solve) is accepted and has nothing to do."""
    B = _lib.as_c128(control_matrix)
    if B.ndim != 3:
        raise ValueError('Expected control_matrix.ndim == 3.')
    n_nops, n_basis, n_omega = B.shape
    phases = _lib.as_c128(np.asarray(phases).reshape(n_omega))
    L = np.asarray(total_propagator_liouville)
    l_complex = np.iscomplexobj(L)
    L = (_lib.as_c128(L) if l_complex else _lib.as_f64(L)).reshape(n_basis, n_basis)
    repeats = int(repeats)
    if repeats < 1:
        raise ValueError(f'Expected repeats >= 1, not {repeats}')
    out = _lib.empty(B.shape)
    if out.size:
        ctx = _lib.context()
        _lib.check(ctx, _lib.lib().ffb_control_matrix_periodic(
            ctx, n_nops, n_basis, n_omega, repeats, _lib.ptr(phases), _lib.ptr(B), _lib.ptr(L),
            int(l_complex), _lib.ptr(out)))
    return out


def _filter_function(control_matrix, which, P):
    B = _lib.as_c128(control_matrix)
    n_nops, n_basis, n_omega = B.shape[-3:]
    gen = which == 'generalized'
    shape = (n_nops, n_nops) + ((n_basis, n_basis) if gen else ()) + (n_omega,)
    if P is not None:
        shape = (P, P) + shape
    F = _lib.empty(shape)
    if F.size:
        ctx = _lib.context()
        _lib.check(ctx, _lib.lib().ffb_filter_function(ctx, 1 if P is None else P, n_nops, n_basis,
                                                       n_omega, _lib.ptr(B), int(gen),
                                                       _lib.ptr(F)))
    if P is not None and gen:
        # kernel layout (g,h,a,b,k,l,o) == reference layout 'ghabklo'
        pass
    return F


@util.parse_optional_parameters(which=('fidelity', 'generalized'))
def calculate_filter_function(control_matrix, which: str = 'fidelity') -> ndarray:
    r"""Filter function :math:`F_{\alpha\beta}(\omega)=\sum_k\tilde{\mathcal B}^*_{\alpha k}
    \tilde{\mathcal B}_{\beta k}` (or the generalized :math:`F_{\alpha\beta,kl}`)."""
    control_matrix = np.asarray(control_matrix)
    if control_matrix.ndim != 3:
        raise ValueError('Expected control_matrix.ndim == 3.')
    return _filter_function(control_matrix, which, None)


@util.parse_optional_parameters(which=('fidelity', 'generalized'))
def calculate_pulse_correlation_filter_function(control_matrix, which: str = 'fidelity') -> ndarray:
    r"""Pulse-correlation filter function :math:`F^{(gg')}_{\alpha\beta}(\omega)` of shape
    (n_pls, n_pls, n_nops, n_nops, [n_basis, n_basis,] n_omega)."""
    control_matrix = np.asarray(control_matrix)
    if control_matrix.ndim != 4:
        raise ValueError('Expected control_matrix.ndim == 4.')
    return _filter_function(control_matrix, which, control_matrix.shape[0])


def _integrate_against_spectrum(filter_function, spectrum, omega, idx, d):
    """integrate(Re(F[..., idx, idx, :] S), omega) / (2 pi d) on the GPU
    (``numeric.py:2318-2320`` with the integrand of ``:259-374``)."""
    omega = _lib.as_f64(omega)
    idx = np.asarray(idx)
    spectrum = util.parse_spectrum(np.asarray(spectrum), omega, idx)
    F = _lib.as_c128(filter_function)
    n_nops, n_omega = F.shape[-2], F.shape[-1]
    lead_shape = F.shape[:-3]
    n_lead = int(np.prod(lead_shape)) if lead_shape else 1
    s_complex = np.iscomplexobj(spectrum)
    S = _lib.as_c128(spectrum) if s_complex else _lib.as_f64(spectrum)
    n_sel = len(idx)
    out_shape = lead_shape + ((n_sel, n_sel) if spectrum.ndim == 3 else (n_sel,))
    out = np.empty(out_shape, dtype=np.float64)
    idx32 = np.ascontiguousarray(idx, dtype=np.int32)
    ctx = _lib.context()
    _lib.check(ctx, _lib.lib().ffb_infidelity(ctx, n_lead, n_nops, n_sel, _lib.ptr(idx32), n_omega,
                                              _lib.ptr(F), _lib.ptr(S), spectrum.ndim,
                                              int(s_complex), _lib.ptr(omega), int(d),
                                              _lib.ptr(out)))
    return out


def _infidelity_convergence(pulse, spectrum, omega, n_oper_identifiers, n_sel, show_progressbar):
    """``test_convergence=True`` (behaviour of reference ``numeric.py:2254-2292``): ``spectrum`` is a
    function of frequency, ``omega`` a dict of grid parameters; the infidelity is evaluated on
    ``n_points`` grids of growing size between ``omega_IR`` and ``omega_UV``.  Returns
    ``(n_samples, infidelities[len(n_samples), n_sel])``."""
    if not callable(spectrum):
        raise TypeError('Spectrum should be callable when test_convergence == True.')
    if not hasattr(omega, 'get'):
        raise TypeError('omega should be dictionary with parameters when test_convergence == True.')
    base = 2*np.pi/pulse.tau
    settings = {'omega_IR': base*1e-2, 'omega_UV': base*1e+2, 'spacing': 'linear', 'n_min': 100,
                'n_max': 500, 'n_points': 10}
    settings.update({key: omega[key] for key in settings if omega.get(key) is not None})
    grids = {'linear': np.linspace, 'log': np.geomspace}
    if settings['spacing'] not in grids:
        raise ValueError("spacing should be either 'linear' or 'log'.")
    make_grid = grids[settings['spacing']]
    step = (settings['n_max'] - settings['n_min'])//(settings['n_points'] - 1)
    n_samples = np.arange(settings['n_min'], settings['n_max'] + step, step)
    results = np.empty((len(n_samples), n_sel))
    for row, n in zip(results, n_samples):
        grid = make_grid(settings['omega_IR'], settings['omega_UV'], n)
        row[:] = infidelity(pulse, spectrum(grid), grid, n_oper_identifiers=n_oper_identifiers,
                            show_progressbar=show_progressbar)
    return n_samples, results


def _smallness_parameter(pulse, spectrum, omega, idx):
    r"""The parameter :math:`\xi` that bounds the validity of the perturbative expansion (reference
    ``numeric.py:2322-2332``): :math:`\xi^2=\sum_\alpha\big[\int\frac{d\omega}{2\pi}S_\alpha(\omega)\big]
    \big[\sum_g s_\alpha^{(g)}\Delta t_g\big]^2\lVert B_\alpha\rVert^2`."""
    if spectrum.ndim > 2:
        raise NotImplementedError('Smallness parameter only implemented for uncorrelated noise '
                                  'sources')
    noise_power = util.integrate(spectrum, np.asarray(omega))/(2*np.pi)
    sensitivity = np.sum(pulse.dt*pulse.n_coeffs[idx], axis=-1)
    operator_norm = util.abs2(pulse.n_opers[idx]).sum(axis=(1, 2))
    return np.sqrt(np.sum(noise_power*sensitivity**2*operator_norm))


@util.parse_optional_parameters(which=('total', 'correlations'))
def infidelity(pulse, spectrum, omega, n_oper_identifiers=None, which: str = 'total',
               show_progressbar: bool = False, cache_intermediates: bool = False,
               return_smallness: bool = False, test_convergence: bool = False):
    r"""Leading-order entanglement infidelity
    :math:`\mathcal I_{\alpha\beta}=\frac1d\int\frac{d\omega}{2\pi}S_{\alpha\beta}(\omega)
    F_{\alpha\beta}(\omega)` for each (pair of) noise operator(s); see the reference docstring
    (``numeric.py:2074-2250``) for the options, which behave identically here."""
    idx = util.get_indices_from_identifiers(pulse.n_oper_identifiers, n_oper_identifiers)

    if test_convergence:
        return _infidelity_convergence(pulse, spectrum, omega, n_oper_identifiers, len(idx),
                                       show_progressbar)

    spectrum = np.asarray(spectrum)
    if (which == 'total' and n_oper_identifiers is None and not cache_intermediates
            and not return_smallness and len(omega) and pulse.basis.istraceless
            and pulse._is_cold()):
        # cold pulse, all noise operators: diagonalisation, control matrix, filter function and the
        # integral in ONE library call (everything stays on the device in between)
        return pulse._cold_pipeline(omega, util.parse_spectrum(spectrum, omega, idx))
    if which == 'total':
        if not pulse.basis.istraceless:
            # Trace tensor enters (numeric.py:2295-2305): F_ab = sum_kl B*_ak T_kl B_bl / d with
            # T_kl = sum_m tr(C_k C_l C_m C_m) - tr(C_k C_m C_l C_m).  Composed from the device ops:
            # B' = B T^T (from_atomic with an empty first pulse), then the (0, 1) block of the
            # pulse-correlation filter function of the pair (B, B').
            traces = pulse.basis.four_element_traces
            T = (np.einsum('klmm->kl', traces) - np.einsum('kmlm->kl', traces))
            B = pulse.get_control_matrix(omega, show_progressbar, cache_intermediates)
            n_omega = B.shape[-1]
            stacked = np.stack([np.zeros_like(B), B])
            Bp = calculate_control_matrix_from_atomic(np.ones((1, n_omega), dtype=complex), stacked,
                                                      np.ascontiguousarray(T.T)[None])
            pair = calculate_pulse_correlation_filter_function(np.stack([B, Bp]))
            filter_function = pair[0, 1]/pulse.d
        else:
            filter_function = pulse.get_filter_function(omega, which='fidelity',
                                                        show_progressbar=show_progressbar,
                                                        cache_intermediates=cache_intermediates)
    else:
        if pulse.is_cached('omega') and not np.array_equal(pulse.omega, omega):
            raise ValueError('Pulse correlation infidelities requested '
                             + 'but omega not equal to cached frequencies.')
        filter_function = pulse.get_pulse_correlation_filter_function()

    infid = _integrate_against_spectrum(filter_function, spectrum, omega, idx, pulse.d)

    if return_smallness:
        return infid, _smallness_parameter(pulse, spectrum, omega, idx)

    return infid


# ------------------------------------------------------------------------------------------------
# decay amplitudes, cumulant function, error transfer matrix
# ------------------------------------------------------------------------------------------------
def _decay_amplitudes_from_control_matrix(control_matrix, spectrum, omega, idx) -> ndarray:
    """trapezoid(Re(conj(B_ak) S_ab B_bl), omega) / (2 pi) on the GPU; the integrand of
    ``numeric.py:344-347`` / ``:362-371`` is never materialised."""
    omega = _lib.as_f64(omega)
    idx = np.asarray(idx)
    spectrum = util.parse_spectrum(np.asarray(spectrum), omega, idx)
    B = _lib.as_c128(control_matrix)
    lead = () if B.ndim == 3 else (B.shape[0], B.shape[0])
    P = 1 if B.ndim == 3 else B.shape[0]
    n_nops, n_basis, n_omega = B.shape[-3:]
    s_complex = np.iscomplexobj(spectrum)
    S = _lib.as_c128(spectrum) if s_complex else _lib.as_f64(spectrum)
    n_sel = len(idx)
    out = np.empty(lead + ((n_sel, n_sel) if spectrum.ndim == 3 else (n_sel,))
                   + (n_basis, n_basis), dtype=np.float64)
    idx32 = np.ascontiguousarray(idx, dtype=np.int32)
    ctx = _lib.context()
    _lib.check(ctx, _lib.lib().ffb_decay_amplitudes(
        ctx, P, n_nops, n_sel, _lib.ptr(idx32), n_basis, n_omega, _lib.ptr(B), _lib.ptr(S),
        spectrum.ndim, int(s_complex), _lib.ptr(omega), _lib.ptr(out)))
    return out


def _decay_amplitudes_from_filter_function(filter_function, spectrum, omega, idx) -> ndarray:
    """Same integral from a cached generalized filter function ([P, P,] a, b, k, l, omega): the basis
    axes are moved in front so that the infidelity kernel integrates every (k, l) row."""
    F = np.asarray(filter_function)
    n_basis = F.shape[-2]
    lead = F.shape[:-5]
    Fm = np.moveaxis(F, (-3, -2), (-5, -4))          # ([P, P,] k, l, a, b, omega)
    out = _integrate_against_spectrum(Fm, spectrum, omega, idx, 1)
    # ([P, P,] k, l, a[, b]) -> ([P, P,] a[, b], k, l)
    n_op_axes = out.ndim - len(lead) - 2
    out = np.moveaxis(out, (len(lead), len(lead) + 1), (-2, -1))
    assert out.shape[-2:] == (n_basis, n_basis) and n_op_axes in (1, 2)
    return np.ascontiguousarray(out)


@util.parse_optional_parameters(which=('total', 'correlations'))
def calculate_decay_amplitudes(pulse, spectrum, omega, n_oper_identifiers=None,
                               which: str = 'total', show_progressbar: bool = False,
                               cache_intermediates: bool = False,
                               memory_parsimonious: bool = False) -> ndarray:
    r"""Decay amplitudes :math:`\Gamma_{\alpha\beta,kl}=\int\frac{d\omega}{2\pi}
    \tilde{\mathcal B}^*_{\alpha k}(\omega)S_{\alpha\beta}(\omega)\tilde{\mathcal B}_{\beta l}
    (\omega)` of shape ([n_pls, n_pls,] n_nops, [n_nops,] n_basis, n_basis) (reference
    ``numeric.py:1194-1337``).  ``memory_parsimonious`` is accepted for compatibility; the kernel never
    builds the integrand, so there is nothing to trade."""
    idx = util.get_indices_from_identifiers(pulse.n_oper_identifiers, n_oper_identifiers)
    if which == 'total':
        if pulse.is_cached('filter_function_gen') and not pulse.is_cached('control_matrix'):
            return _decay_amplitudes_from_filter_function(
                pulse.get_filter_function(omega, which='generalized'), spectrum, omega, idx)
        control_matrix = pulse.get_control_matrix(omega, show_progressbar, cache_intermediates)
    else:
        if pulse.is_cached('omega') and not np.array_equal(pulse.omega, omega):
            raise ValueError('Pulse correlation decay amplitudes requested but omega not '
                             + 'equal to cached frequencies.')
        if (pulse.is_cached('filter_function_pc_gen')
                and not pulse.is_cached('control_matrix_pc')):
            return _decay_amplitudes_from_filter_function(
                pulse.get_pulse_correlation_filter_function(which='generalized'), spectrum, omega,
                idx)
        control_matrix = pulse.get_pulse_correlation_control_matrix()
    return _decay_amplitudes_from_control_matrix(control_matrix, spectrum, omega, idx)


def _contract_with_trace_tensor(decay_amplitudes, basis) -> ndarray:
    r""":math:`\sum_{kl}\Gamma_{kl}(T_{klji}-T_{kjli}-T_{kilj}+T_{kijl})` with
    :math:`T_{ijkl}=\mathrm{tr}(C_iC_jC_kC_l)` (reference ``numeric.py:1160-1165``) WITHOUT the
    n_basis^4 trace tensor: with :math:`D_k=\sum_l\Gamma_{kl}C_l`, :math:`G=\sum_kC_kD_k`,
    :math:`G'=\sum_kD_kC_k` and the map :math:`\Phi(X)=\sum_kC_kXD_k`,

    .. math:: \mathrm{tr}(GC_jC_i)-\mathrm{tr}(\Phi(C_j)C_i)-\mathrm{tr}(\Phi(C_i)C_j)
              +\mathrm{tr}(G'C_iC_j).

    O(n_basis^2 d^2 + n_basis d^4) numbers instead of n_basis^4 (69 GB for d = 16)."""
    C = np.asarray(basis)
    Gamma = np.asarray(decay_amplitudes)
    D = np.einsum('...kl,lab->...kab', Gamma, C)
    G1 = np.einsum('kab,...kbc->...ac', C, D)
    G2 = np.einsum('...kab,kbc->...ac', D, C)
    t1 = np.einsum('...jab,iba->...ij', np.einsum('...ac,jcb->...jab', G1, C), C)
    t4 = np.einsum('...iab,jba->...ij', np.einsum('...ac,icb->...iab', G2, C), C)
    Phi = np.einsum('kab,...kcd->...adbc', C, D)            # Phi(X)[a,d] = Phi[a,d,b,c] X[b,c]
    PhiC = np.einsum('...adbc,jbc->...jad', Phi, C)
    t2 = np.einsum('...jad,ida->...ij', PhiC, C)
    return t1 - t2 - t2.swapaxes(-1, -2) + t4


@util.parse_optional_parameters(which=('total', 'correlations'))
def calculate_cumulant_function(pulse, spectrum=None, omega=None, n_oper_identifiers=None,
                                which: str = 'total', second_order: bool = False,
                                decay_amplitudes: Optional[ndarray] = None,
                                frequency_shifts: Optional[ndarray] = None,
                                show_progressbar: bool = False, memory_parsimonious: bool = False,
                                cache_intermediates: Optional[bool] = None) -> ndarray:
    r"""Cumulant function :math:`\mathcal K(\tau)` of shape ([[n_pls, n_pls,] n_nops,] n_nops,
    n_basis, n_basis) (reference ``numeric.py:957-1191``); first order only."""
    if second_order or frequency_shifts is not None:
        raise NotImplementedError('Second-order (frequency shift) terms are out of scope of '
                                  'filter_functions_b200 (SURVEY.md section 2, row 11)')
    N, d = pulse.basis.shape[:2]
    if spectrum is None and omega is None and decay_amplitudes is None:
        raise ValueError('Require either spectrum and frequencies or precomputed '
                         + 'decay amplitudes (frequency shifts)')
    if decay_amplitudes is None:
        decay_amplitudes = calculate_decay_amplitudes(pulse, spectrum, omega, n_oper_identifiers,
                                                      which, show_progressbar,
                                                      bool(cache_intermediates),
                                                      memory_parsimonious)
    decay_amplitudes = np.asarray(decay_amplitudes)

    if d == 2 and pulse.basis.btype in ('Pauli', 'GGM'):
        # single qubit (numeric.py:1120-1143): K_ij = Gamma_ij off the diagonal,
        # K_ii = -sum_{k != i, k > 0} Gamma_kk; the identity row / column vanishes.
        K = np.zeros(decay_amplitudes.shape, decay_amplitudes.dtype)
        off = np.zeros((N, N), dtype=bool)
        off[1:, 1:] = ~np.eye(N - 1, dtype=bool)
        K[..., off] = decay_amplitudes[..., off]
        diag = np.einsum('...ii->...i', decay_amplitudes)[..., 1:]
        for i in range(1, N):
            K[..., i, i] = -(diag.sum(axis=-1) - diag[..., i - 1])
        return K

    return -0.5*_contract_with_trace_tensor(decay_amplitudes, pulse.basis).real


def error_transfer_matrix(pulse=None, spectrum=None, omega=None, n_oper_identifiers=None,
                          second_order: bool = False, cumulant_function: Optional[ndarray] = None,
                          show_progressbar: bool = False, memory_parsimonious: bool = False,
                          cache_intermediates: bool = False) -> ndarray:
    r"""Error transfer matrix :math:`\langle\tilde{\mathcal U}\rangle=\exp\mathcal K(\tau)` of
    shape (n_basis, n_basis), summed over all noise operators (reference ``numeric.py:1938-2059``)."""
    from scipy import linalg as sla
    if cumulant_function is None:
        if any(arg is None for arg in (pulse, spectrum, omega)):
            raise ValueError('Require either precomputed cumulant function '
                             + 'or pulse, spectrum, and omega as arguments.')
        cumulant_function = calculate_cumulant_function(
            pulse, spectrum, omega, n_oper_identifiers, 'total', second_order,
            show_progressbar=show_progressbar, memory_parsimonious=memory_parsimonious,
            cache_intermediates=cache_intermediates)
    if not hasattr(cumulant_function, 'sum'):
        raise TypeError(f'cumulant_function invalid type: {type(cumulant_function)}')
    if np.ndim(cumulant_function) < 2:
        raise ValueError(f'cumulant_function invalid shape: {np.shape(cumulant_function)}')
    # all leading axes enumerate noise operators (pairs): their cumulants add up in the exponent
    summed = cumulant_function.reshape(-1, *cumulant_function.shape[-2:]).sum(axis=0)
    if summed.shape[0] != summed.shape[1]:
        raise ValueError(f'cumulant_function invalid shape: {cumulant_function.shape}')
    return sla.expm(summed)
