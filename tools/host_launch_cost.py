"""Host-side cost of enqueueing one device-resident step (GPU box): is the front of the step launch-bound?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import workloads
from filter_functions_b200.device import DevicePulse

wl = workloads.get(sys.argv[1] if len(sys.argv) > 1 else 'c2')
dev = DevicePulse(wl.c_opers, wl.c_coeffs, wl.n_opers, wl.n_coeffs, wl.dt, wl.basis, wl.omega, wl.spectrum, device=0)
dev.bind_stream()
for _ in range(5):
    dev.step()
torch.cuda.synchronize()
names = ['diagonalize', 'calculate_control_matrix', 'calculate_filter_function', 'calculate_infidelity']
acc = {n: 0.0 for n in names}
N = 200
for _ in range(N):
    torch.cuda.synchronize()
    for n in names:
        t0 = time.perf_counter()
        getattr(dev, n)()
        acc[n] += time.perf_counter() - t0
torch.cuda.synchronize()
for n in names:
    print('%-28s %.1f us host time to enqueue' % (n, acc[n]/N*1e6))
# device time of the front (diagonalize only) and of a full step with the CPU far ahead
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize()
big = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
big.zero_(); big.zero_()           # ~0.35 ms of GPU work: the CPU gets ahead
e[0].record(); dev.diagonalize(); e[1].record(); dev.calculate_control_matrix(); e[2].record()
dev.calculate_filter_function(); dev.calculate_infidelity(); e[3].record()
torch.cuda.synchronize()
print('GPU time with the CPU ahead: diagonalize %.1f us, control matrix %.1f us, ff+infidelity %.1f us'
      % (e[0].elapsed_time(e[1])*1e3, e[1].elapsed_time(e[2])*1e3, e[2].elapsed_time(e[3])*1e3))
