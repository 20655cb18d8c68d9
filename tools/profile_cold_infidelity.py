"""cProfile of a cold ff.infidelity(pulse, S, omega) call (config 2): where the host time goes (GPU box)."""
import cProfile, os, pstats, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
import filter_functions_b200 as ff

name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
wl = workloads.get(name)
pulse = ff.PulseSequence([[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
                         [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
                         wl.dt, ff.Basis.pauli(int(np.log2(wl.d))))


def step():
    pulse.cleanup('all')
    return ff.infidelity(pulse, wl.spectrum, wl.omega)


for _ in range(5):
    step()
t0 = time.perf_counter()
for _ in range(50):
    step()
print('%s: %.3f ms per cold call' % (name, (time.perf_counter() - t0)/50*1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    step()
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
