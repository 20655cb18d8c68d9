"""cProfile of one config-5 step (cache 9 gate control matrices + concatenate) -- host-side overheads."""
import cProfile, pstats, sys, os, io
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
import filter_functions_b200 as ff
n_omega = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
omega = np.logspace(-2, 2, n_omega)
pulses = workloads.build_qft_pulses(ff, 4)
def step():
    for p in pulses:
        p.cleanup('frequency dependent')
        p.cache_control_matrix(omega)
    return ff.concatenate(pulses, omega=omega)
for _ in range(3): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(35); print(s.getvalue()[:6000])
