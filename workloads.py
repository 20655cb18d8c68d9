"""Synthetic inputs of the BASELINE.json configurations, as plain NumPy
arrays so that the engine, the oracle and the tests all consume the same data.  No network, no
datasets: everything is drawn from ``numpy.random.default_rng(seed)``.
"""
from dataclasses import dataclass, field

import numpy as np

I2, X, Y, Z = (np.array(m, dtype=complex) for m in
               ([[1, 0], [0, 1]], [[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]))


def kron(*ops):
    out = np.array([[1.0 + 0j]])
    for op in ops:
        out = np.kron(out, op)
    return out


def pauli_basis(n):
    elems = np.array([I2, X, Y, Z])
    out = elems
    for _ in range(n - 1):
        out = np.einsum('aij,bkl->abikjl', out, elems).reshape(len(out)*4, out.shape[1]*2,
                                                               out.shape[2]*2)
    return out/np.sqrt(2**n)


@dataclass
class Workload:
    name: str
    description: str
    c_opers: np.ndarray
    c_ids: list
    c_coeffs: np.ndarray
    n_opers: np.ndarray
    n_ids: list
    n_coeffs: np.ndarray
    dt: np.ndarray
    basis: np.ndarray
    basis_kind: str
    omega: np.ndarray
    spectrum: np.ndarray
    extra: dict = field(default_factory=dict)

    @property
    def d(self):
        return self.c_opers.shape[-1]

    @property
    def G(self):
        return len(self.dt)

    @property
    def t(self):
        return np.concatenate(([0.0], self.dt.cumsum()))

    @property
    def flops_per_seg_omega(self):
        """W = 8 n_nops n_basis d^2 + 12 d^2."""
        d = self.d
        return 8*len(self.n_opers)*len(self.basis)*d*d + 12*d*d

    @property
    def trig_per_seg_omega(self):
        """T = 2 d^2 + 2 sin/cos evaluations in the reference formulation."""
        return 2*self.d**2 + 2

    def with_omega(self, omega, spectrum=None):
        out = Workload(**{**self.__dict__})
        out.omega = np.asarray(omega)
        out.spectrum = self.spectrum_fn(out.omega) if spectrum is None else spectrum
        return out

    def spectrum_fn(self, omega):
        return self.extra['spectrum_fn'](omega)


def readme_hadamard():
    """C1: README.md:17-33 of the reference (QuTiP objects replaced by arrays)."""
    dt = np.array([1.0, 1.0])
    tau = dt.sum()
    omega = np.geomspace(2*np.pi*1e-2/tau, 2*np.pi*10/dt.min(), 300)
    fn = lambda w: 1e-2/w  # noqa: E731
    return Workload('c1', 'README Hadamard: Y/2 then X, sigma_z noise, 300 log-spaced frequencies',
                    np.array([X/2, Y/2]), ['X', 'Y'], np.array([[0, np.pi], [np.pi/2, 0]]),
                    np.array([Z/2]), ['Z'], np.ones((1, 2)), dt, pauli_basis(1), 'pauli', omega,
                    fn(omega), {'spectrum_fn': fn, 'expected_infidelity': 0.00253303})


def single_qubit_grape(G=10_000, n_omega=10_000, seed=2):
    """C2: single-qubit random GRAPE-style pulse, 3 Pauli noise operators, 1/f spectrum."""
    rng = np.random.default_rng(seed)
    dt = np.full(G, 0.05)
    tau = G*0.05
    omega = np.geomspace(2*np.pi*1e-2/tau, 2*np.pi*10/0.05, n_omega)
    fn = lambda w: 1e-2/w  # noqa: E731
    return Workload('c2', f'single-qubit GRAPE-style pulse: d=2, G={G}, n_nops=3, Pauli basis (4), '
                    f'n_omega={n_omega}, 1/f spectrum',
                    np.array([X/2, Y/2]), ['X', 'Y'], rng.standard_normal((2, G))*np.pi,
                    np.array([X/2, Y/2, Z/2]), ['X', 'Y', 'Z'], np.ones((3, G)), dt,
                    pauli_basis(1), 'pauli', omega, fn(omega), {'spectrum_fn': fn})


def two_qubit_exchange(G=2000, n_omega=50_000, seed=3):
    """C3: exchange-coupled singlet-triplet two-qubit gate, 6 noise operators."""
    rng = np.random.default_rng(seed)
    dt = np.full(G, 0.2)
    tau = dt.sum()
    eps = np.cumsum(rng.normal(0, 0.05, (3, G)), axis=1) - 1
    J = np.exp(eps)
    ZI, IZ, ZZ = kron(Z, I2), kron(I2, Z), kron(Z, Z)
    XI, IX = kron(X, I2), kron(I2, X)
    c_opers = np.array([ZI/2, IZ/2, ZZ/4, XI/2, IX/2])
    c_coeffs = np.vstack([J, np.full((1, G), 0.1), np.full((1, G), 0.7)])
    n_opers = np.array([ZI/2, IZ/2, ZZ/4, XI/2, IX/2, (XI + IX)/4])
    n_coeffs = np.vstack([J, np.ones((3, G))])
    omega = np.geomspace(1/tau, 1e2, n_omega)
    eps0 = 2.7241e-4
    alpha = 0.7
    A = 4e-11/eps0**2*(2*np.pi*1e-3)**alpha
    fn = lambda w: A/w**alpha  # noqa: E731
    return Workload('c3' if G == 2000 else 'd4', f'two-qubit exchange gate: d=4, G={G}, n_nops=6, '
                    f'Pauli basis (16), n_omega={n_omega}, 1/f^0.7 spectrum',
                    c_opers, ['ZI', 'IZ', 'ZZ', 'XI', 'IX'], c_coeffs,
                    n_opers, ['eps1', 'eps2', 'eps3', 'bx1', 'bx2', 'bxx'], n_coeffs, dt,
                    pauli_basis(2), 'pauli', omega, fn(omega), {'spectrum_fn': fn})


def north_star_d4(G=10_000, n_omega=10_000, seed=3):
    """north_star target shape: d=4, 1e4 segments, 1e4 frequencies (operators of C3)."""
    wl = two_qubit_exchange(G, n_omega, seed)
    wl.name = 'd4'
    return wl


def get(name, **kwargs):
    return {'c1': readme_hadamard, 'c2': single_qubit_grape, 'c3': two_qubit_exchange,
            'd4': north_star_d4}[name](**kwargs)
