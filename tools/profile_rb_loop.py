"""cProfile of the per-call ff.concatenate + ff.infidelity loop of config 4 (GPU box)."""
import cProfile, os, pstats, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads
import filter_functions_b200 as ff

omega = workloads.rb_omega()
S = workloads.rb_spectrum(omega)
cliffords = workloads.build_cliffords(ff, omega)
rows = workloads.rb_sequences(200)

def loop(n):
    return [ff.infidelity(ff.concatenate([cliffords[k] for k in row]), S, omega) for row in rows[:n]]
loop(20)
t0 = time.perf_counter(); loop(100); t = time.perf_counter() - t0
print('per sequence: %.3f ms' % (t*10))
t0 = time.perf_counter(); [ff.concatenate([cliffords[k] for k in row]) for row in rows[:100]]; t = time.perf_counter() - t0
print('concatenate only: %.3f ms' % (t*10))
pr = cProfile.Profile(); pr.enable(); loop(100); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
