"""FFB_TRACE timeline of the fused pipeline for the README pulse (config 1)."""
import os, sys, time
os.environ['FFB_TRACE'] = '1'
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import filter_functions_b200 as ff
from filter_functions_b200 import util
X, Y, Z = util.paulis[1:]
pulse = ff.PulseSequence([[X/2, [0, np.pi], 'X'], [Y/2, [np.pi/2, 0], 'Y']], [[Z/2, [1, 1], 'Z']], [1, 1])
omega = util.get_sample_frequencies(pulse)
S = 1e-2/omega
for i in range(8):
    pulse.cleanup('all')
    t0 = time.perf_counter()
    ff.infidelity(pulse, S, omega)
    print('call %d: %.0f us wall' % (i, (time.perf_counter() - t0)*1e6), file=sys.stderr)
