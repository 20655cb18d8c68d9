"""Where does the deviation of the d4 control matrix from the reference's full-size fixture sit?
(GPU box)  Prints the worst frequencies / rows of the FP64 path (and of the int8 path with --int8)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if '--int8' in sys.argv:
    os.environ['FFB_CTRLMAT_INT8'] = '1'
import workloads
import filter_functions_b200 as ff

name = 'd4'
wl = workloads.get(name)
g = np.load(os.path.join(ROOT, 'tests', 'golden', f'workload_full_{name}.npz'))
pulse = ff.PulseSequence([[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
                         [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
                         wl.dt, ff.Basis.pauli(2))
B = pulse.get_control_matrix(wl.omega)
pick = g['pick']
err = np.abs(B[:, :, pick] - g['control_matrix'])/g['scale'][:, None, None]      # (n_nops, n_basis, 1024)
print('worst overall %.3e' % err.max())
per_w = err.max(axis=(0, 1))
worst = np.argsort(per_w)[::-1][:12]
for i in worst:
    j, k = np.unravel_index(err[:, :, i].argmax(), err.shape[:2])
    print('omega[%5d] = %.6e  err %.2e  at (nop %d, basis %d)  |B|/scale %.2e' % (
        pick[i], wl.omega[pick[i]], per_w[i], j, k, abs(B[j, k, pick[i]])/g['scale'][j]))
print('median over frequencies %.2e, 90th pct %.2e' % (np.median(per_w), np.quantile(per_w, 0.9)))
print('per basis element (max over nops, freq):', np.array2string(err.max(axis=(0, 2)), precision=1))
print('propagators vs fixture total: %.2e' % np.abs(pulse.total_propagator - g['total_propagator']).max())
lo = per_w[wl.omega[pick] < 1.0].max() if (wl.omega[pick] < 1.0).any() else 0
print('max err for omega < 1: %.2e ; for omega > 10: %.2e' % (lo, per_w[wl.omega[pick] > 10].max()))
