"""FFB_TRACE timeline of numeric.calculate_control_matrix_from_scratch from host arrays (GPU box)."""
import os, sys, time
os.environ['FFB_TRACE'] = '1'
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
import filter_functions_b200 as ff
from filter_functions_b200 import numeric

for name in sys.argv[1:] or ['c2', 'c3']:
    wl = workloads.get(name)
    pulse = ff.PulseSequence([[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
                             [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
                             wl.dt, ff.Basis.pauli(int(np.log2(wl.d))))
    pulse.diagonalize()
    ev, V, Q = pulse.eigvals, pulse.eigvecs, pulse.propagators
    for i in range(5):
        t0 = time.perf_counter()
        B = numeric.calculate_control_matrix_from_scratch(ev, V, Q, wl.omega, pulse.basis, pulse.n_opers,
                                                          pulse.n_coeffs, pulse.dt, pulse.t)
        print(name, 'call %d: %.0f us wall' % (i, (time.perf_counter() - t0)*1e6), file=sys.stderr)
