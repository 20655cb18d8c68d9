"""Accuracy of the GPU diagonalisation + propagator scan against LAPACK / sequential NumPy (GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import workloads
import ff_oracle as oracle
import filter_functions_b200 as ff

for name in sys.argv[1:] or ['d4', 'c2']:
    wl = workloads.get(name)
    H = oracle.hamiltonian_from_coeffs(wl.c_opers, wl.c_coeffs)
    ev_o, V_o, Q_o = oracle.diagonalize(H, wl.dt)
    ev, V, Q = ff.numeric.diagonalize(H, wl.dt)
    G, d = ev.shape
    print(name, 'G', G, 'd', d)
    print('  eigenvalues: max |dE| %.2e (||H|| ~ %.2f)' % (np.abs(ev - ev_o).max(), np.abs(ev_o).max()))
    defect = np.abs(np.einsum('gji,gjk->gik', V.conj(), V) - np.eye(d)).max(axis=(1, 2))
    defect_o = np.abs(np.einsum('gji,gjk->gik', V_o.conj(), V_o) - np.eye(d)).max(axis=(1, 2))
    print('  unitarity defect of V: gpu max %.2e mean %.2e | lapack max %.2e mean %.2e' % (defect.max(), defect.mean(), defect_o.max(), defect_o.mean()))
    tr = np.einsum('gji,gji->g', V.conj(), V).real/d - 1
    tr_o = np.einsum('gji,gji->g', V_o.conj(), V_o).real/d - 1
    print('  mean of (||V||_F^2/d - 1): gpu %.2e  lapack %.2e   (a bias accumulates linearly in the product)' % (tr.mean(), tr_o.mean()))
    P = np.einsum('gij,gkj->gik', Q[1:], Q[:-1].conj())        # P_g = Q_{g+1} Q_g^dagger (if Q unitary)
    P_o = np.einsum('gij,gkj->gik', Q_o[1:], Q_o[:-1].conj())
    print('  per-segment propagator (from consecutive Q): max diff %.2e' % np.abs(P - P_o).max())
    print('  cumulative propagators: max |Q - Q_ref| at g = G/4, G/2, G: %.2e %.2e %.2e' % tuple(
        np.abs(Q[g] - Q_o[g]).max() for g in (G//4, G//2, G)))
    for label, QQ in (('gpu', Q), ('numpy', Q_o)):
        dd = [np.abs(QQ[g].conj().T @ QQ[g] - np.eye(d)).max() for g in (G//4, G//2, G)]
        print('  unitarity defect of Q (%s) at G/4, G/2, G: %.2e %.2e %.2e' % ((label,) + tuple(dd)))
    # the same chain multiplied sequentially on the host from the GPU's own eigensystems
    Pg = np.einsum('gij,gj,gkj->gik', V, np.exp(-1j*ev*wl.dt[:, None]), V.conj())
    Qs = np.empty_like(Q); Qs[0] = np.eye(d)
    for g in range(G):
        Qs[g + 1] = Pg[g] @ Qs[g]
    print('  sequential host product of the GPU eigensystems vs reference: %.2e ; vs GPU scan: %.2e' % (
        np.abs(Qs[G] - Q_o[G]).max(), np.abs(Qs[G] - Q[G]).max()))
