"""Latency of ff.infidelity on small pulses (BASELINE config 1 and a mid-size pulse), cold cache."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
which = sys.argv[1] if len(sys.argv) > 1 else 'b200'
if which == 'reference':
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim')); sys.path.insert(0, '/root/reference')
    import filter_functions as ff
    from filter_functions import util
else:
    sys.path.insert(0, ROOT)
    import filter_functions_b200 as ff
    from filter_functions_b200 import util
X, Y, Z = util.paulis[1:]
def timeit(f, n=200):
    for _ in range(10): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0)/n*1e6
pulse = ff.PulseSequence([[X/2, [0, np.pi], 'X'], [Y/2, [np.pi/2, 0], 'Y']], [[Z/2, [1, 1], 'Z']], [1, 1])
omega = util.get_sample_frequencies(pulse)
S = 1e-2/omega
def c1():
    pulse.cleanup('all'); return ff.infidelity(pulse, S, omega)
print(which, 'config 1 (2 segments x 300 omega): %.0f us per cold ff.infidelity' % timeit(c1), c1())
rng = np.random.default_rng(0)
G = 100
p2 = ff.PulseSequence([[X/2, rng.standard_normal(G), 'X'], [Y/2, rng.standard_normal(G), 'Y']],
                      [[Z/2, np.ones(G), 'Z'], [X/2, np.ones(G), 'Xn']], np.full(G, 0.1))
om2 = np.geomspace(1e-2, 1e2, 500)
S2 = 1/om2
def mid():
    p2.cleanup('all'); return ff.infidelity(p2, S2, om2)
print(which, 'mid (100 segments x 500 omega): %.0f us per cold ff.infidelity' % timeit(mid, 100))
