"""Where does the end-to-end time of ff.infidelity(pulse, S, omega) go?  (run on the GPU box)"""
import sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
import filter_functions_b200 as ff
from filter_functions_b200 import numeric, util, superoperator

def bench(f, n=10):
    f(); f()
    t0 = time.perf_counter()
    for _ in range(n): r = f()
    return (time.perf_counter() - t0)/n*1e3

for name in sys.argv[1:] or ['c2', 'c3']:
    wl = workloads.get(name)
    pulse = ff.PulseSequence([[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
                             [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
                             wl.dt, ff.Basis.pauli(int(np.log2(wl.d))))
    def full():
        pulse.cleanup('all'); return ff.infidelity(pulse, wl.spectrum, wl.omega)
    def cold():
        pulse.cleanup('all'); pulse.omega = wl.omega; pulse._cold_pipeline(wl.omega)
    print(name, 'full infidelity        %.3f ms' % bench(full))
    print(name, 'cold pipeline (+cache) %.3f ms' % bench(cold))
    F = pulse.get_filter_function(wl.omega); B = pulse.get_control_matrix(wl.omega)
    print(name, 'integrate vs spectrum  %.3f ms' % bench(lambda: numeric._integrate_against_spectrum(F, wl.spectrum, wl.omega, np.arange(len(wl.n_opers)), wl.d)))
    print(name, 'liouville              %.3f ms' % bench(lambda: superoperator.liouville_representation(pulse.total_propagator, pulse.basis)))
    print(name, 'cexp total phases      %.3f ms' % bench(lambda: util.cexp(pulse.omega*pulse.tau)))
    print(name, 'filter function (host) %.3f ms' % bench(lambda: numeric.calculate_filter_function(B)))
    print(name, 'diagonalize (host)     %.3f ms' % bench(lambda: numeric._diagonalize_from_coeffs(pulse.c_opers, pulse.c_coeffs, pulse.dt)))
    ev, V, Q = pulse.eigvals, pulse.eigvecs, pulse.propagators
    print(name, 'control matrix (host)  %.3f ms' % bench(lambda: numeric.calculate_control_matrix_from_scratch(ev, V, Q, wl.omega, pulse.basis, pulse.n_opers, pulse.n_coeffs, pulse.dt, pulse.t)))
    print(name, 'omega setter           %.3f ms' % bench(lambda: setattr(pulse, 'omega', wl.omega)))
    big = np.empty_like(B)
    print(name, 'np.empty+copy of B     %.3f ms  (%.1f MB)' % (bench(lambda: np.copyto(big, B)), B.nbytes/1e6))
