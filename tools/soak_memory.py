"""Soak test: thousands of calls through the public API; device and host memory must stay flat (GPU box)."""
import gc, os, resource, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import workloads
import filter_functions_b200 as ff


def snapshot(tag):
    free, total = torch.cuda.mem_get_info()
    rss = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss/1024
    print(f'{tag:32s} device used {(total - free)/2**20:9.1f} MiB   host max RSS {rss:9.1f} MiB', flush=True)
    return (total - free)/2**20, rss


omega = workloads.rb_omega()
S = workloads.rb_spectrum(omega)
cliffords = workloads.build_cliffords(ff, omega)
rows = workloads.rb_sequences(1000)
wl = workloads.get('c2')
pulse = ff.PulseSequence([[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
                         [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
                         wl.dt, ff.Basis.pauli(1))
snapshot('start')
marks = []
for rnd in range(4):
    t0 = time.perf_counter()
    for row in rows:
        ff.infidelity(ff.concatenate([cliffords[k] for k in row]), S, omega)
    for _ in range(100):
        pulse.cleanup('all')
        ff.infidelity(pulse, wl.spectrum, wl.omega)
    for _ in range(20):      # fresh pulse objects: caches keyed by object identity must not accumulate
        p2 = ff.PulseSequence([[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
                              [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
                              wl.dt, ff.Basis.pauli(1))
        ff.infidelity(p2, wl.spectrum, wl.omega)
        q = p2 @ p2
        del p2, q
    gc.collect()
    marks.append(snapshot(f'after round {rnd} ({time.perf_counter() - t0:.1f} s)'))
grow_dev = marks[-1][0] - marks[1][0]
grow_host = marks[-1][1] - marks[1][1]
print(f'growth between round 1 and round 3: device {grow_dev:.1f} MiB, host {grow_host:.1f} MiB')
print('SOAK_OK' if grow_dev < 64 and grow_host < 64 else 'SOAK_GROWTH')
