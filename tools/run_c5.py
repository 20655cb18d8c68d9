#!/usr/bin/env python
"""Config 5 (4-qubit QFT from cached gate pulses) once or a few times -- target for ncu launch lists /
captures of the concatenation kernels:  python tools/run_c5.py [n_omega] [repeats]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import workloads  # noqa: E402
import filter_functions_b200 as ff  # noqa: E402

n_omega = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 1
omega = np.logspace(-2, 2, n_omega)
pulses = workloads.build_qft_pulses(ff, 4)
for p in pulses:
    p.cache_control_matrix(omega)
for _ in range(repeats):
    t0 = time.perf_counter()
    qft = ff.concatenate(pulses, omega=omega)
    print(f'concatenate: {time.perf_counter() - t0:.4f} s', flush=True)
