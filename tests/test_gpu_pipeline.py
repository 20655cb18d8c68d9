"""The fused cold-cache pipeline behind ``ff.infidelity`` / ``get_filter_function``
(``ffb_pulse_filter_function``): packed upload, frequency blocks with overlapped strided downloads,
``t`` formed in the library, and the per-stage ``dt`` test of the thread-per-frequency kernel.  All
compared with the oracle on the same inputs (rtol 1e-10 as normalised max error)."""
import numpy as np
import pytest

import ff_oracle as oracle
from helpers import nerr, rand_herm_traceless

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pulse_and_reference(ff, rng, d, G, n_nops, dt, omega, spectrum):
    c_opers = rand_herm_traceless(rng, d, 2)
    n_opers = rand_herm_traceless(rng, d, n_nops)
    c_coeffs = rng.standard_normal((2, G))
    n_coeffs = rng.random((n_nops, G)) + 0.5
    basis = ff.Basis.pauli(int(np.log2(d))) if d in (2, 4) else ff.Basis.ggm(d)
    ids = [f'n{i}' for i in range(n_nops)]
    pulse = ff.PulseSequence([[o, c, f'c{i}'] for i, (o, c) in enumerate(zip(c_opers, c_coeffs))],
                             [[o, c, i] for o, c, i in zip(n_opers, n_coeffs, ids)], dt, basis)
    H = oracle.hamiltonian_from_coeffs(c_opers, c_coeffs)
    ev, V, Q = oracle.diagonalize(H, dt)
    B = oracle.control_matrix_from_scratch(ev, V, Q, omega, np.asarray(basis), n_opers, n_coeffs, dt)
    F = oracle.filter_function(B)
    return pulse, Q, B, F


@pytest.mark.parametrize('n_blocks', [1, 2, 3, 5])
@pytest.mark.parametrize('d,n_nops,n_omega', [(2, 3, 4500), (4, 2, 4133)])
def test_frequency_blocks(engine, monkeypatch, n_blocks, d, n_nops, n_omega):
    """Any number of frequency blocks gives the oracle's result (block edges are multiples of 64
    frequencies; 4133 and 4500 are not), through both entry points that use the fused call."""
    ff = engine
    rng = np.random.default_rng(100*d + n_blocks)
    G = 37
    dt = 1 - rng.random(G)
    omega = np.geomspace(1e-2, 50, n_omega)
    spectrum = 1/omega
    pulse, Q, B, F = _pulse_and_reference(ff, rng, d, G, n_nops, dt, omega, spectrum)
    monkeypatch.setenv('FFB_PIPELINE_BLOCKS', str(n_blocks))
    infid = ff.infidelity(pulse, spectrum, omega)
    assert nerr(pulse.propagators, Q) < TOL
    assert nerr(pulse.get_control_matrix(omega), B) < TOL
    F_gpu = pulse.get_filter_function(omega)
    assert nerr(F_gpu, F) < TOL
    # trapezoid of the oracle's filter function (util.integrate, util.py:880-906)
    f = np.einsum('aaw->aw', F).real*spectrum
    ref = 0.5*((f[:, 1:] + f[:, :-1])*np.diff(omega)).sum(-1)/(2*np.pi*d)
    assert nerr(infid, ref) < TOL
    # and the blocked result is the unblocked one up to summation order
    monkeypatch.setenv('FFB_PIPELINE_BLOCKS', '1')
    pulse.cleanup('all')
    assert nerr(pulse.get_filter_function(omega), F_gpu) < 1e-13


def test_time_axis_formed_in_library(engine):
    """``t`` is not computed on the host for a cold pulse; the library's [0, cumsum(dt)] must be the
    one NumPy gives (bitwise: the phase argument omega*t_g is rounded as in the reference)."""
    ff = engine
    rng = np.random.default_rng(5)
    G, d, n_nops = 300, 2, 3
    dt = 1 - rng.random(G)
    omega = np.geomspace(1e-1, 1e3, 200)
    pulse, Q, B, F = _pulse_and_reference(ff, rng, d, G, n_nops, dt, omega, 1/omega)
    assert 't' not in pulse._data
    pulse.get_filter_function(omega)
    B_lib_t = pulse.get_control_matrix(omega).copy()
    assert 't' not in pulse._data            # still not needed on the host
    pulse.cleanup('all')
    _ = pulse.t                              # host copy cached -> it is passed down instead
    pulse.get_filter_function(omega)
    assert np.array_equal(pulse.get_control_matrix(omega), B_lib_t)
    assert nerr(B_lib_t, B) < TOL


@pytest.mark.parametrize('d,n_nops,G', [(2, 3, 8*7 + 5), (3, 1, 43), (2, 1, 17), (2, 3, 250),
                                        (3, 1, 333)])
def test_piecewise_uniform_time_grid(engine, d, n_nops, G):
    """Thread-per-frequency kernel: the half-angle factors are refreshed per stage of 8 segments
    only when a dt of the stage differs.  Grids that are uniform, then irregular, then uniform again
    (with run lengths that are not multiples of the stage; G = 250, 333 give every warp several
    stages) exercise both loops and the hand-over between them."""
    ff = engine
    rng = np.random.default_rng(17*d + G)
    dt = np.full(G, 0.3)
    a, b, c = G//3 + 1, G//3 + 9, 2*G//3 + 3
    dt[a:b] = 1 - rng.random(b - a)
    dt[b:c] = 0.7
    dt[-2] = 0.05
    omega = np.concatenate(([0.0], np.geomspace(1e-3, 80, 333)))
    pulse, Q, B, F = _pulse_and_reference(ff, rng, d, G, n_nops, dt, omega, np.ones_like(omega))
    assert nerr(pulse.get_control_matrix(omega), B) < TOL
    assert nerr(pulse.get_filter_function(omega), F) < TOL


def _call_pipeline(ff, wl_arrays, outs, spectrum):
    """ffb_pulse_filter_function through ctypes with caller-owned arrays (``outs``: name -> array or None)."""
    from filter_functions_b200 import _lib
    c_opers, c_coeffs, n_opers, n_coeffs, dt, basis, omega = wl_arrays
    G, d = len(dt), c_opers.shape[-1]
    p = _lib.ptr
    ctx = _lib.context()
    S = np.ascontiguousarray(spectrum, dtype=np.float64)
    _lib.check(ctx, _lib.lib().ffb_pulse_filter_function(
        ctx, G, d, len(c_opers), len(n_opers), len(basis), len(omega), p(c_opers), p(c_coeffs),
        p(n_opers), p(n_coeffs), p(dt), None, p(basis), p(omega), p(S), S.ndim, 0,
        p(outs.get('eigvals')), p(outs.get('eigvecs')), p(outs.get('propagators')),
        p(outs.get('control_matrix')), p(outs.get('filter_function')), p(outs.get('infidelity')),
        p(outs.get('total_phases')), p(outs.get('liouville'))))


@pytest.mark.parametrize('layout', ['separate', 'one_numpy_block', 'partial'])
def test_pipeline_result_layouts(engine, layout):
    """The C entry point with result arrays that are NOT one block of the library's pool: separate
    pageable NumPy arrays, views into one ordinary NumPy buffer (adjacent, but not the library's:
    nothing between them may be touched), and some results not requested (NULL)."""
    ff = engine
    rng = np.random.default_rng(11)
    d, G, n_nops, n_omega = 2, 29, 3, 400
    dt = 1 - rng.random(G)
    omega = np.geomspace(1e-2, 30, n_omega)
    spectrum = 1/omega
    pulse, Q, B, F = _pulse_and_reference(ff, rng, d, G, n_nops, dt, omega, spectrum)
    arrays = (np.ascontiguousarray(pulse.c_opers), np.ascontiguousarray(pulse.c_coeffs),
              np.ascontiguousarray(pulse.n_opers), np.ascontiguousarray(pulse.n_coeffs),
              np.ascontiguousarray(pulse.dt), np.ascontiguousarray(np.asarray(pulse.basis)), omega)
    n_basis = 4
    shapes = dict(eigvals=((G, d), np.float64), eigvecs=((G, d, d), np.complex128),
                  propagators=((G + 1, d, d), np.complex128),
                  total_phases=((n_omega,), np.complex128), liouville=((n_basis, n_basis), np.complex128),
                  control_matrix=((n_nops, n_basis, n_omega), np.complex128),
                  filter_function=((n_nops, n_nops, n_omega), np.complex128),
                  infidelity=((n_nops,), np.float64))
    guard = None
    if layout == 'one_numpy_block':
        # views 8 bytes apart in one buffer; the guard words in between must survive
        sizes = {k: int(np.prod(sh))*np.dtype(dt_).itemsize for k, (sh, dt_) in shapes.items()}
        buf = np.full(sum(sizes.values()) + 8*(len(sizes) + 1), 0x5A, dtype=np.uint8)
        outs, off, guard = {}, 8, []
        for k, (sh, dt_) in shapes.items():
            outs[k] = buf[off:off + sizes[k]].view(dt_).reshape(sh)
            guard.append((off + sizes[k], off + sizes[k] + 8))
            off += sizes[k] + 8
    else:
        outs = {k: np.empty(sh, dt_) for k, (sh, dt_) in shapes.items()}
    if layout == 'partial':
        for k in ('eigvecs', 'total_phases', 'control_matrix'):
            outs[k] = None
    _call_pipeline(ff, arrays, outs, spectrum)
    assert nerr(outs['propagators'], Q) < TOL
    assert nerr(outs['filter_function'], F) < TOL
    if outs['control_matrix'] is not None:
        assert nerr(outs['control_matrix'], B) < TOL
    f = np.einsum('aaw->aw', F).real*spectrum
    ref = 0.5*((f[:, 1:] + f[:, :-1])*np.diff(omega)).sum(-1)/(2*np.pi*d)
    assert nerr(outs['infidelity'], ref) < TOL
    if guard is not None:
        assert (buf[:8] == 0x5A).all() and all((buf[a:b] == 0x5A).all() for a, b in guard)
