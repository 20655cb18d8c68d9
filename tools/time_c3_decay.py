"""Config 3 extras: decay amplitudes and error transfer matrix on the cached control matrix (GPU box)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
import filter_functions_b200 as ff
from filter_functions_b200 import numeric

wl = workloads.get('c3')
pulse = ff.PulseSequence([[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
                         [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
                         wl.dt, ff.Basis.pauli(2))
pulse.cache_filter_function(wl.omega)
def best(f, n=5):
    f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return min(ts)*1e3, r
t_dec, Gamma = best(lambda: numeric.calculate_decay_amplitudes(pulse, wl.spectrum, wl.omega))
t_cum, K = best(lambda: numeric.calculate_cumulant_function(pulse, wl.spectrum, wl.omega))
t_etm, U = best(lambda: numeric.error_transfer_matrix(pulse, wl.spectrum, wl.omega))
print('c3 decay amplitudes %.3f ms %s, cumulant function %.3f ms, error transfer matrix %.3f ms %s'
      % (t_dec, Gamma.shape, t_cum, t_etm, U.shape))
