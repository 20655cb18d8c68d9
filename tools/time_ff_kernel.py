"""Times ``ffb_dev_filter_function`` (fidelity) on device-resident control matrices of several shapes,
with the one-pass Gram kernel (default for n_basis >= 32) and with the row-pair kernel
(``FFB_FF_GRAM=0``), and reports the fraction of the measured HBM rate the algorithmic bytes
``16 (L n_basis + L^2) n_omega`` reach.  Run as two processes (the switch is read once per process):

    python tools/time_ff_kernel.py; FFB_FF_GRAM=0 python tools/time_ff_kernel.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from filter_functions_b200 import _lib  # noqa: E402

SHAPES = [(1, 18, 256, 10000),    # config 5
          (1, 6, 64, 20000), (1, 12, 64, 20000), (1, 16, 64, 20000), (1, 36, 64, 10000),
          (9, 18, 256, 1000),     # pulse correlations of config 5 (162 rows)
          (1, 6, 16, 50000)]      # config 3 (row-pair kernel either way)


def main():
    dev = torch.device('cuda', 0)
    ctx = _lib.context(0)
    L = _lib.lib()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(ctx, L.ffb_set_stream(ctx, stream, 1))
    peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gram = os.environ.get('FFB_FF_GRAM', '1') != '0'
    for P, n_nops, n_basis, n_omega in SHAPES:
        g = torch.Generator(device=dev).manual_seed(1)
        B = torch.randn((P, n_nops, n_basis, n_omega, 2), dtype=torch.float64, device=dev, generator=g)
        rows = P*n_nops
        F = torch.empty((rows, rows, n_omega, 2), dtype=torch.float64, device=dev)
        times = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(ctx, L.ffb_dev_filter_function(ctx, P, n_nops, n_basis, n_omega, B.data_ptr(), 0,
                                                      F.data_ptr()))
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[3:]))
        nbytes = 16.0*(rows*n_basis + rows*rows)*n_omega
        flops = 8.0*(rows*(rows + 1)/2)*n_basis*n_omega
        print(json.dumps({'kernel': 'gram' if gram and n_basis >= 32 and rows >= 4 else 'row-pair',
                          'P': P, 'n_nops': n_nops, 'n_basis': n_basis, 'n_omega': n_omega,
                          'ms': round(ms, 4), 'algorithmic_MB': round(nbytes/1e6, 1),
                          'GBps': round(nbytes/ms/1e6, 1), 'hbm_frac': round(nbytes/ms/1e6/peak, 3),
                          'min_TFLOPs': round(flops/ms/1e9, 2)}))
        del B, F


if __name__ == '__main__':
    main()
