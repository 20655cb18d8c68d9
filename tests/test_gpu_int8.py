"""The fixed-point control-matrix kernel on the int8 tensor cores (csrc/ffb_ctrlmat_i8.cu, opt-in through
FFB_CTRLMAT_INT8=1, d = 4): parity with the oracle / the reference fixtures inside the north-star
tolerance (1e-10 normalised per noise operator) and agreement with the FP64 kernels."""
import os

import numpy as np
import pytest

import ff_oracle as oracle
import workloads
from helpers import nerr, rand_herm

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TOL = 1e-10


class int8_path:
    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        self.old = os.environ.get('FFB_CTRLMAT_INT8')
        os.environ['FFB_CTRLMAT_INT8'] = '1' if self.on else '0'

    def __exit__(self, *exc):
        if self.old is None:
            os.environ.pop('FFB_CTRLMAT_INT8', None)
        else:
            os.environ['FFB_CTRLMAT_INT8'] = self.old


@pytest.mark.parametrize('G,n_nops,n_omega', [(64, 6, 64), (67, 3, 130), (300, 6, 257), (1000, 5, 100),
                                              (2050, 6, 64), (4100, 1, 33)])
def test_int8_control_matrix_against_oracle(engine, G, n_nops, n_omega):
    ff = engine
    rng = np.random.default_rng(G + 17*n_nops)
    d = 4
    c_opers = rand_herm(rng, d, 2)
    n_opers = rand_herm(rng, d, n_nops)
    c_coeffs = rng.standard_normal((2, G))
    n_coeffs = rng.random((n_nops, G)) + 0.5
    dt = np.where(rng.random(G) < 0.5, 0.3, 0.7) if G == 67 else np.full(G, 0.4)
    H = oracle.hamiltonian_from_coeffs(c_opers, c_coeffs)
    ev, V, Q = oracle.diagonalize(H, dt)
    basis = oracle.pauli_basis(2)
    # includes omega = 0, a negative frequency and frequencies next to level splittings
    gaps = np.abs(ev[G//2][:, None] - ev[G//2][None, :])
    near = gaps[gaps > 0][:4]*(1 + 1e-9)
    omega = np.concatenate(([0.0, -0.37], near, np.geomspace(1e-3, 40, n_omega - 2 - len(near))))
    t = np.concatenate(([0], dt.cumsum()))
    with int8_path(False):
        B_f64 = ff.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs,
                                                                 dt, t)
    with int8_path(True):
        B_i8 = ff.numeric.calculate_control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs,
                                                                dt, t)
    B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega, basis, n_opers, n_coeffs, dt, t)
    assert np.isfinite(B_i8.view(float)).all()
    for j in range(n_nops):
        assert nerr(B_f64[j], B_o[j]) < 1e-13
        assert nerr(B_i8[j], B_o[j]) < TOL
    # the fixed-point path is a different kernel, not the FP64 one in disguise
    assert not np.array_equal(B_i8, B_f64)


@pytest.mark.parametrize('name', ['c3', 'd4'])
def test_int8_full_size_workloads(engine, name):
    """BASELINE config 3 and the north-star shape through the public API with the int8 path enabled,
    against the reference's own results (tests/golden/workload_full_*.npz)."""
    ff = engine
    wl = workloads.get(name)
    g = np.load(os.path.join(GOLDEN, f'workload_full_{name}.npz'))
    with int8_path(True):
        pulse = ff.PulseSequence(
            [[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
            [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)], wl.dt, ff.Basis.pauli(2))
        infid = ff.infidelity(pulse, wl.spectrum, wl.omega)
        B = pulse.get_control_matrix(wl.omega)
        F = pulse.get_filter_function(wl.omega)
    gp = g['pick']
    worst = 0.0
    for j in range(len(wl.n_opers)):
        err = np.abs(B[j][:, gp] - g['control_matrix'][j]).max()/g['scale'][j]
        worst = max(worst, err)
        assert err < TOL
    assert nerr(F[..., gp], g['filter_function']) < TOL
    np.testing.assert_allclose(infid, g['infidelity'], rtol=TOL)
    print(f'{name}: int8 path, worst normalised control-matrix error {worst:.2e}')
