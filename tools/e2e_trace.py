"""Host-side timeline of the fused pipeline behind ff.infidelity (run on the GPU box): FFB_TRACE=1."""
import os, sys, time
os.environ['FFB_TRACE'] = '1'
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
import filter_functions_b200 as ff

for name in sys.argv[1:] or ['c2']:
    wl = workloads.get(name)
    pulse = ff.PulseSequence([[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
                             [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
                             wl.dt, ff.Basis.pauli(int(np.log2(wl.d))))
    for i in range(6):
        pulse.cleanup('all')
        t0 = time.perf_counter()
        ff.infidelity(pulse, wl.spectrum, wl.omega)
        print(name, 'call %d: %.0f us wall' % (i, (time.perf_counter() - t0)*1e6), file=sys.stderr)
