"""CPU emulation of the int8 digit scheme for the control-matrix GEMM: B = A (rows x K, real) @ P (K x n_w, complex),
K = 13 G.  5 balanced base-256 digits per operand, products with i + j >= 4 kept, exact integer accumulation."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import workloads, ff_oracle as oracle

name, G, n_w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
wl = workloads.get(name)
sl = slice(0, G)
order = np.argsort(wl.n_ids)
H = oracle.hamiltonian_from_coeffs(wl.c_opers, wl.c_coeffs[:, sl])
dt = wl.dt[sl]
ev, V, Q = oracle.diagonalize(H, dt)
pick = np.unique(np.concatenate([np.arange(0, len(wl.omega), len(wl.omega)//n_w)]))[:n_w]
omega = wl.omega[pick]
t = np.concatenate(([0], dt.cumsum()))
B_ref = oracle.control_matrix_from_scratch(ev, V, Q, omega, wl.basis, wl.n_opers[order], wl.n_coeffs[order][:, sl], dt, t)
d = wl.d
n_nops, n_basis = len(wl.n_opers), len(wl.basis)
# transformed operators
Bbar = np.einsum('gba,jbc,gcd->gjad', V.conj(), wl.n_opers[order], V)*wl.n_coeffs[order][:, sl].T[:, :, None, None]
U = np.einsum('gba,gbc->gac', Q[:-1].conj(), V)          # Q_g^+ V_g
Cbar = np.einsum('gba,kbc,gcd->gkad', U.conj(), wl.basis, U)
pairs = [(m, n) for m in range(d) for n in range(m + 1, d)]
rows = n_nops*n_basis
A = np.zeros((rows, G, 13))
diag = np.einsum('gjmm,gkmm->gjk', Bbar.real, Cbar.real)
A[:, :, 0] = diag.reshape(G, rows).T
for p, (m, n) in enumerate(pairs):
    prod = Bbar[:, :, None, m, n]*Cbar[:, None, :, n, m]
    A[:, :, 1 + 2*p] = prod.real.reshape(G, rows).T
    A[:, :, 2 + 2*p] = prod.imag.reshape(G, rows).T
def I(x, dtg):
    with np.errstate(all='ignore'):
        return np.where(x == 0, dtg, (np.exp(1j*x*dtg) - 1)/(1j*np.where(x == 0, 1, x)))
P = np.zeros((G, 13, len(omega)), dtype=complex)
for g in range(G):
    ph = np.exp(1j*omega*t[g])
    P[g, 0] = ph*I(omega, dt[g])
    for p, (m, n) in enumerate(pairs):
        Om = ev[g, m] - ev[g, n]
        jp, jm = I(omega + Om, dt[g]), I(omega - Om, dt[g])
        P[g, 1 + 2*p] = ph*(jp + jm)
        P[g, 2 + 2*p] = ph*1j*(jp - jm)
B_f64 = np.einsum('rgk,gkw->rw', A, P).reshape(n_nops, n_basis, -1)
scale = np.abs(B_ref).max(axis=(1, 2))
print('f64 reformulation vs oracle:', max(np.abs(B_f64[j] - B_ref[j]).max()/scale[j] for j in range(n_nops)))
# ---- quantise
D = 5
def digits(q):
    C = sum(128*256**i for i in range(D))
    u = (q + C).astype(np.int64)
    return [(((u >> (8*i)) & 0xFF) - 128).astype(np.int64) for i in range(D)]
sA = 2.0**np.ceil(np.log2(np.abs(A).max(axis=(1, 2)).clip(1e-300)))
sP = 2.0**np.ceil(np.log2(2*dt.max()))
qA = np.rint(A/sA[:, None, None]*2.0**38).astype(np.int64)
qPr = np.rint(P.real/sP*2.0**38).astype(np.int64); qPi = np.rint(P.imag/sP*2.0**38).astype(np.int64)
dA = digits(qA); dPr = digits(qPr); dPi = digits(qPi)
assert all(np.abs(x).max() <= 128 for x in dA + dPr + dPi)
res = np.zeros((rows, len(omega)), dtype=complex)
for tlev in range(4, 9):
    acc_r = np.zeros((rows, len(omega)), dtype=np.int64); acc_i = acc_r.copy()
    for i in range(D):
        j = tlev - i
        if 0 <= j < D:
            a = dA[j].reshape(rows, -1)      # coefficient digit j
            acc_r += a @ dPr[i].reshape(-1, len(omega))
            acc_i += a @ dPi[i].reshape(-1, len(omega))
    print('level', tlev, 'max |acc| = 2^%.1f' % np.log2(max(np.abs(acc_r).max(), np.abs(acc_i).max(), 1)))
    res += (acc_r + 1j*acc_i)*256.0**tlev
res *= (sA[:, None]*sP*2.0**-76)
B_i8 = res.reshape(n_nops, n_basis, -1)
print('int8 emulation vs oracle   :', [float(np.abs(B_i8[j] - B_ref[j]).max()/scale[j]) for j in range(n_nops)])
print('int8 emulation vs f64 form :', [float(np.abs(B_i8[j] - B_f64[j]).max()/scale[j]) for j in range(n_nops)])
print('scale', scale, 'sA range', sA.min(), sA.max(), 'sP', sP)
