"""Times the from-scratch control matrix on device-resident random pulses of shapes outside the bench
workloads (d = 3, 8, 16), and checks a few frequencies against the oracle.  (GPU box)

    python tools/time_ctrlmat_shapes.py; FFB_CTRLMAT_STATIC=0 python tools/time_ctrlmat_shapes.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch  # noqa: E402

import ff_oracle as oracle  # noqa: E402
from filter_functions_b200.device import DevicePulse  # noqa: E402

SHAPES = [(4, 4000, 5, 8000), (8, 2000, 3, 5000), (8, 2000, 1, 5000), (8, 500, 6, 2000), (3, 4000, 3, 8000), (16, 100, 2, 2000)]


def herm(rng, d, n):
    a = rng.standard_normal((n, d, d)) + 1j*rng.standard_normal((n, d, d))
    return a + a.conj().swapaxes(1, 2)


def main():
    rng = np.random.default_rng(5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for d, G, n_nops, n_omega in SHAPES:
        c_opers, n_opers = herm(rng, d, 2), herm(rng, d, n_nops)
        c_coeffs, n_coeffs = rng.standard_normal((2, G)), rng.random((n_nops, G)) + 0.5
        dt = np.full(G, 0.3)
        basis = oracle.ggm_basis(d)
        omega = np.geomspace(1e-3, 30, n_omega)
        dev = DevicePulse(c_opers, c_coeffs, n_opers, n_coeffs, dt, basis, omega)
        dev.bind_stream()
        dev.diagonalize()
        times = []
        for _ in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dev.calculate_control_matrix()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[2:]))
        pick = np.unique(np.linspace(0, n_omega - 1, 5).astype(int))
        B = dev.control_matrix.cpu().numpy()
        ev, V, Q = (x.cpu().numpy() for x in (dev.eigvals, dev.eigvecs, dev.propagators))
        B_o = oracle.control_matrix_from_scratch(ev, V, Q, omega[pick], basis, n_opers, n_coeffs, dt)
        err = max(np.abs(B[j][:, pick] - B_o[j]).max()/np.abs(B[j]).max() for j in range(n_nops))
        W = 8*n_nops*d**4 + 12*d*d
        print(json.dumps({'d': d, 'G': G, 'n_nops': n_nops, 'rows': n_nops*d*d, 'n_omega': n_omega,
                          'static': os.environ.get('FFB_CTRLMAT_STATIC', '1') != '0', 'ms': round(ms, 3),
                          'seg_omega_per_s': G*n_omega/ms*1e3, 'reference_formulation_TFLOPs': W*G*n_omega/ms*1e-9,
                          'err_vs_oracle': err}))


if __name__ == '__main__':
    main()
