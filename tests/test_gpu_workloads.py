"""The BASELINE.json workloads on the GPU: reduced sizes against the reference's fixtures, and the
full bench sizes through size-independent properties plus oracle spot checks (the oracle's cost is
linear in n_omega, so a few frequencies at full G take seconds)."""
import os

import numpy as np
import pytest

import ff_oracle as oracle
import workloads
from helpers import nerr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 1e-10


def make_pulse(ff, wl):
    basis = ff.Basis.pauli(int(np.log2(wl.d)))
    return ff.PulseSequence(
        [[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
        [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)], wl.dt, basis)


@pytest.mark.parametrize('name,kwargs', [('c2', dict(G=64, n_omega=96)),
                                         ('c3', dict(G=40, n_omega=64))])
def test_reduced_workloads_against_reference(engine, name, kwargs):
    g = np.load(os.path.join(GOLDEN, 'workloads_small.npz'))
    wl = workloads.get(name, **kwargs)
    pulse = make_pulse(engine, wl)
    assert list(pulse.n_oper_identifiers) == list(g[f'{name}_n_ids'])
    assert nerr(pulse.get_control_matrix(wl.omega), g[f'{name}_control_matrix']) < TOL
    assert nerr(pulse.get_filter_function(wl.omega), g[f'{name}_filter_function']) < TOL
    assert nerr(engine.infidelity(pulse, wl.spectrum, wl.omega), g[f'{name}_infidelity']) < TOL


@pytest.mark.parametrize('name', ['c2', 'c3', 'd4'])
def test_full_size_workloads(engine, name):
    """Full bench size: EVERY frequency the reference fixture holds (tests/golden/workload_full_*.npz,
    computed once by the unmodified reference: the whole grid for config 2; 1024 frequencies -- both
    grid ends, the neighbourhoods of the level splittings where the kernels take their fix-up branch,
    a random rest -- plus the infidelities of the whole grid for the d = 4 shapes), an oracle spot
    check on 12 more frequencies, device-resident path == API path,
    filter function Hermitian and positive on the diagonal, infidelity additive over frequency
    blocks (the property the omega-sharded multi-GPU path relies on)."""
    ff = engine
    wl = workloads.get(name)
    pulse = make_pulse(ff, wl)
    F = pulse.get_filter_function(wl.omega)
    B = pulse.get_control_matrix(wl.omega)
    n_nops = len(wl.n_opers)
    assert B.shape == (n_nops, len(wl.basis), len(wl.omega))
    assert np.isfinite(B.view(float)).all()
    # spot check against the oracle
    order = np.argsort(wl.n_ids)
    pick = np.unique(np.linspace(0, len(wl.omega) - 1, 12).astype(int))
    H = oracle.hamiltonian_from_coeffs(wl.c_opers, wl.c_coeffs)
    ev, V, Q = oracle.diagonalize(H, wl.dt)
    assert nerr(pulse.eigvals, ev) < TOL and nerr(pulse.propagators, Q) < TOL
    B_o = oracle.control_matrix_from_scratch(ev, V, Q, wl.omega[pick], wl.basis, wl.n_opers[order],
                                             wl.n_coeffs[order], wl.dt, wl.t)
    scale = np.abs(B).max(axis=(1, 2))
    for j in range(n_nops):
        assert np.abs(B[j][:, pick] - B_o[j]).max() < TOL*scale[j]
    assert nerr(F[..., pick], oracle.filter_function(B_o)) < TOL
    # the reference's own results at full size
    g = np.load(os.path.join(GOLDEN, f'workload_full_{name}.npz'))
    assert list(pulse.n_oper_identifiers) == list(g['n_ids'])
    gp = g['pick']
    assert len(gp) == (len(wl.omega) if name == 'c2' else 1024)
    for j in range(n_nops):     # normalised per noise operator, scale = max over the WHOLE grid
        assert np.abs(B[j][:, gp] - g['control_matrix'][j]).max() < TOL*g['scale'][j]
        assert abs(scale[j] - g['scale'][j]) < TOL*g['scale'][j]
    assert nerr(F[..., gp], g['filter_function']) < TOL
    assert nerr(pulse.total_propagator, g['total_propagator']) < TOL
    # structure
    assert nerr(F, F.conj().transpose(1, 0, 2)) < 1e-14
    assert (F[range(n_nops), range(n_nops)].real >= 0).all()
    # infidelity: whole grid == sum over two halves sharing one abscissa
    full = ff.infidelity(pulse, wl.spectrum, wl.omega)
    mid = len(wl.omega)//2
    halves = 0
    for sl in (slice(0, mid + 1), slice(mid, None)):
        part = make_pulse(ff, wl)
        halves = halves + ff.infidelity(part, wl.spectrum[sl], wl.omega[sl])
    np.testing.assert_allclose(halves, full, rtol=1e-10)
    np.testing.assert_allclose(full, g['infidelity'], rtol=TOL)
    want = oracle.infidelity_from_filter_function(F, wl.spectrum, wl.omega, wl.d)
    np.testing.assert_allclose(full, want, rtol=1e-12)
    # device-resident path (what bench.py times) gives the same numbers
    from filter_functions_b200.device import DevicePulse
    dev = DevicePulse(wl.c_opers[np.argsort(wl.c_ids)], wl.c_coeffs[np.argsort(wl.c_ids)],
                      wl.n_opers[order], wl.n_coeffs[order], wl.dt, wl.basis, wl.omega, wl.spectrum)
    dev.bind_stream()
    dev.step()
    import torch
    torch.cuda.synchronize()
    assert nerr(dev.control_matrix.cpu().numpy(), B) < 1e-13
    np.testing.assert_allclose(dev.infidelity.cpu().numpy(), full, rtol=1e-12)
    from filter_functions_b200 import _lib
    _lib.check(dev.ctx, _lib.lib().ffb_set_stream(dev.ctx, None, 0))
    # the step-by-step entry points from host arrays (large eigensystems go up straight from the
    # caller's arrays, small inputs through the packed staging block) give the same control matrix
    ev2, V2, Q2 = ff.numeric.diagonalize(H, wl.dt)
    assert nerr(Q2, Q) < TOL
    sub = slice(0, len(wl.omega), max(1, len(wl.omega)//700))
    B_sub = ff.numeric.calculate_control_matrix_from_scratch(
        pulse.eigvals, pulse.eigvecs, pulse.propagators, wl.omega[sub], wl.basis, wl.n_opers[order],
        wl.n_coeffs[order], wl.dt, wl.t)
    assert nerr(B_sub, B[..., sub]) < 1e-12
