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


# ------------------------------------------------------------------------------------------------
# Workloads defined through the public API (BASELINE.json configs 4 and 5).  ``ff`` is the package to
# build them with: ``filter_functions_b200`` in the tests and the bench, the reference itself in
# ``oracle/gen_golden.py`` -- the same code drives both.
# ------------------------------------------------------------------------------------------------
CLIFFORD_WORDS = [                    # examples/randomized_benchmarking.py:158-183 (X = X2, Y = Y2)
    'YYYY', 'XX', 'YY', 'YYXX', 'XY', 'XYYY', 'XXXY', 'XXXYYY', 'YX', 'YXXX', 'YYYX', 'YYYXXX',
    'X', 'XXX', 'Y', 'YYY', 'XYYYXXX', 'XXXYYYX', 'XXY', 'XXYYY', 'YYX', 'YYXXX', 'XYX', 'XYYYX']


def rb_omega(T=20.0, m_max=151, n_omega=301):
    """examples/randomized_benchmarking.py:145."""
    return np.geomspace(1e-2/(7*m_max*T), 1e2/T, n_omega)*2*np.pi


def rb_spectrum(omega, alpha=0.7):
    """examples/randomized_benchmarking.py:189-194."""
    eps0 = 2.7241e-4
    return 4e-11*(2*np.pi*1e-3/omega)**alpha/eps0**2


def build_cliffords(ff, omega, T=20.0):
    """C4: the 24 single-qubit Cliffords from 'naive' X/2 and Y/2 gates with sigma_x noise
    (examples/randomized_benchmarking.py:95-111, :148-183), control matrices cached on ``omega``."""
    X2 = ff.PulseSequence([[X/2, [np.pi/2/T], 'X']], [[X/2, [1], 'X']], [T])
    Y2 = ff.PulseSequence([[Y/2, [np.pi/2/T], 'Y']], [[X/2, [1], 'X']], [T])
    X2.cache_control_matrix(omega)
    Y2.cache_control_matrix(omega)
    gates = {'X': X2, 'Y': Y2}
    cliffords = []
    for word in CLIFFORD_WORDS:
        pulse = gates[word[0]]
        for letter in word[1:]:
            pulse = pulse @ gates[letter]
        cliffords.append(pulse)
    return cliffords


def rb_sequences(n_seq=1000, length=100, seed=4):
    return np.random.default_rng(seed).integers(0, 24, (n_seq, length))


def _embed(op, k, N):
    """op on qubit k of N (k = 0 is the left-most tensor factor)."""
    return kron(*[op if i == k else I2 for i in range(N)])


def build_qft_pulses(ff, N=4, tau=1.0):
    """C5: the gate pulses of the N-qubit QFT of examples/qft.py:42-136 with QuTiP objects replaced by
    Kronecker products: T_I, [H_k, P_{k+1}] for k < N-1, H_{N-1}, T_F."""
    d = 2**N

    def ident(letters):
        out = ['I']*N
        for pos, letter in letters:
            out[pos] = letter
        return ''.join(out)

    def R_k(k, theta, phi):
        Xk, Yk = _embed(X, k, N), _embed(Y, k, N)
        H_c = [[Xk, [theta/2/tau*np.cos(phi)], ident([(k, 'X')])],
               [Yk, [theta/2/tau*np.sin(phi)], ident([(k, 'Y')])]]
        H_n = [[Xk/np.sqrt(d), [1], ident([(k, 'X')])], [Yk/np.sqrt(d), [1], ident([(k, 'Y')])]]
        return ff.PulseSequence(H_c, H_n, [tau])

    def T_pulse(coeff):
        H_c = [[_embed(Z, k - 1, N), [coeff(k)], ident([(k - 1, 'Z')])] for k in range(1, N + 1)]
        H_n = [[_embed(Z, k - 1, N)/np.sqrt(d), [1], ident([(k - 1, 'Z')])]
               for k in range(1, N + 1)]
        return ff.PulseSequence(H_c, H_n, [tau])

    def P_n(n):
        H_c, H_n = [], []
        for l in range(n + 1, N + 1):
            ZZ = _embed(Z, n - 1, N) @ _embed(Z, l - 1, N)
            name = ident([(n - 1, 'Z'), (l - 1, 'Z')])
            H_c.append([ZZ, [-np.pi/4*2**(n - l)/tau], name])
            H_n.append([ZZ/np.sqrt(d), [1], name])
        return ff.PulseSequence(H_c, H_n, [tau])

    def H_k(k):
        return ff.concatenate([R_k(k, np.pi, 0), R_k(k, np.pi/2, -np.pi/2)])

    pulses = [T_pulse(lambda k: np.pi/4*(1 - 2**(1 - k))/tau)]
    for n in range(N - 1):
        pulses.append(H_k(n))
        pulses.append(P_n(n + 1))
    pulses.append(H_k(N - 1))
    pulses.append(T_pulse(lambda k: np.pi/4*(1 - 2**(k - N))/tau))
    return pulses


def qft_matrix(N):
    d = 2**N
    j, k = np.meshgrid(np.arange(d), np.arange(d))
    return np.exp(2j*np.pi*j*k/d)/np.sqrt(d)


def bit_reversal(N):
    d = 2**N
    perm = [int(format(i, f'0{N}b')[::-1], 2) for i in range(d)]
    return np.eye(d)[perm]
