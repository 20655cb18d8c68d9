#!/usr/bin/env python
"""Benchmark of the hot path: from-scratch control matrix -> filter function -> infidelity.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|d4] [--impl reference]

A *step* is one pass of the whole path over one synthetic pulse: diagonalise the
G-segment control Hamiltonian, build the first-order control matrix on n_omega frequencies, reduce it
to the fidelity filter function and integrate against the noise spectrum.  The metric is
BASELINE.json's: control-matrix throughput in segment*frequency pairs per second (fp64).

  value : whole-job seg*omega/s with all inputs resident in HBM (DevicePulse, ffb_dev_* entry points),
          device-timed with CUDA events, L2 flushed between steps, max over ranks.
  e2e   : the same metric through the public NumPy API (ff.infidelity on a cold PulseSequence), host
          buffers in, host arrays out, copies inside the timed region.
  N > 1 : the frequency axis is sharded over the ranks (weak scaling: n_omega per GPU is fixed), the
          per-segment operands are replicated, one NCCL all-reduce of the partial infidelities.
  --impl reference : the reference algorithm on the host cores (the NumPy oracle port of
          oracle/ff_oracle.py, pinned against the reference; the reference itself is Python and its
          dependencies opt_einsum/sparse are not installable here), bounded sample per step.
"""
import argparse
import json
import os
import sys
import threading
import time

if '--impl' in sys.argv and 'reference' in sys.argv:
    # torchrun pins OMP_NUM_THREADS=1 for its workers; the reference arm is the reference's CPU path
    # on ALL host cores, so undo that before NumPy/OpenBLAS load.
    for _var in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_var] = str(os.cpu_count() or 1)

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line to
# stdout at init, nvcc output of a first build, ...): keep a private handle on the real stdout for the
# result line and point file descriptor 1 at stderr for everything else, native code included.
sys.stdout.flush()
RESULT_OUT = os.fdopen(os.dup(1), 'w')
os.dup2(2, 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import workloads  # noqa: E402

METRIC = 'control-matrix throughput (segments*omega/s, fp64), control matrix + filter function + infidelity'
UNIT = 'seg*omega/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--workload', default='c2', choices=['c2', 'c3', 'd4'])
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-seconds', type=float, default=12.0,
                    help='target duration of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    return ap.parse_args()


def workload_config(wl, n_gpus, per_gpu_omega):
    return {
        'workload': f'{wl.name}: {wl.description}',
        'G': wl.G, 'd': wl.d, 'n_nops': len(wl.n_opers), 'n_basis': len(wl.basis),
        'n_omega_per_gpu': per_gpu_omega, 'n_omega_total': per_gpu_omega*n_gpus,
        'parallelism': f'omega-sharded x{n_gpus}' if n_gpus > 1 else 'single GPU',
        'l2': 'flushed between timed steps (256 MiB device write)',
    }


# ------------------------------------------------------------------------------------------------
# CPU leg (oracle port of the reference algorithm)
# ------------------------------------------------------------------------------------------------
def cpu_sample(wl, seconds):
    """Time the reference algorithm on a bounded sample: the first G_cpu segments of the same pulse
    (cost is exactly linear in G, numeric.py:846) on all frequencies, plus filter function and
    infidelity.  Returns (seg*omega/s, description, cores)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ff_oracle as oracle
    H = oracle.hamiltonian_from_coeffs(wl.c_opers, wl.c_coeffs)
    t0 = time.perf_counter()
    ev, V, Q = oracle.diagonalize(H, wl.dt)
    t_diag = time.perf_counter() - t0
    # calibrate on a few segments
    n_cal = 8
    t0 = time.perf_counter()
    oracle.control_matrix_from_scratch(ev, V, Q, wl.omega, wl.basis, wl.n_opers, wl.n_coeffs,
                                       wl.dt, wl.t, segments=range(n_cal))
    per_seg = (time.perf_counter() - t0)/n_cal
    G_cpu = int(max(16, min(wl.G, seconds/per_seg)))
    t0 = time.perf_counter()
    B = oracle.control_matrix_from_scratch(ev, V, Q, wl.omega, wl.basis, wl.n_opers, wl.n_coeffs,
                                           wl.dt, wl.t, segments=range(G_cpu))
    F = oracle.filter_function(B)
    oracle.infidelity_from_filter_function(F, wl.spectrum, wl.omega, wl.d)
    t_ctrl = time.perf_counter() - t0
    total = t_ctrl + t_diag*G_cpu/wl.G
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([p.get('num_threads', 1) for p in threadpool_info()] or [1])
    except Exception:
        blas_threads = os.cpu_count()
    sample = (f'first {G_cpu} of {wl.G} segments x {len(wl.omega)} frequencies (+ filter function, '
              f'infidelity, pro-rata diagonalisation), {total:.1f} s, NumPy/OpenBLAS with '
              f'{blas_threads} BLAS threads on {os.cpu_count()} host cores')
    return G_cpu*len(wl.omega)/total, sample, blas_threads


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    wl = workloads.get(args.workload)
    per_step = max(1.0, min(args.cpu_seconds, 150.0/max(1, args.steps + args.warmup)))
    vals = []
    sample, cores = '', 1
    for i in range(args.warmup + args.steps):
        v, sample, cores = cpu_sample(wl, per_step)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals)) if vals else 0.0
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3*wl.G*len(wl.omega)/value if value else None,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(wl, args.gpus, len(wl.omega)),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
               0x4: 'sw_power_cap', 0x80: 'hw_power_brake_slowdown'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [],
                    'note': 'NVML unavailable'}
        return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.build()
    import filter_functions_b200 as ff
    from filter_functions_b200 import _lib
    from filter_functions_b200 import distributed as ffd
    from filter_functions_b200.device import DevicePulse

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch N>1 with: python -m torch.distributed.run --nnodes=1 '
                             f'--nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py '
                             f'--gpus {args.gpus} ...')
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    base = workloads.get(args.workload)
    n_per = len(base.omega)
    # weak scaling: the global grid has world * n_per frequencies over the same band
    lo, hi = base.omega[0], base.omega[-1]
    omega_global = np.geomspace(lo, hi, n_per*world)
    start, stop = ffd.frequency_shard(len(omega_global), rank, world)
    wl = base.with_omega(omega_global[start:stop])
    G, n_omega_local = wl.G, len(wl.omega)

    dev = DevicePulse(wl.c_opers, wl.c_coeffs, wl.n_opers, wl.n_coeffs, wl.dt, wl.basis, wl.omega,
                      wl.spectrum, device=local_rank)
    ctx = dev.ctx
    L = _lib.lib()
    stream = torch.cuda.current_stream()
    dev.bind_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    total = torch.zeros_like(dev.infidelity)

    def step():
        dev.step()
        if world > 1:
            total.copy_(dev.infidelity)
            dist.all_reduce(total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()

    # ---- device-resident timing -------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    _lib.check(ctx, L.ffb_kernel_timing_enable(ctx, 1))
    import ctypes
    _lib.check(ctx, L.ffb_kernel_timing_read(ctx, None, None, 1))
    launches0 = L.ffb_launch_count(ctx)
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
    barrier()
    for e0, e1 in events:
        flush.zero_()
        e0.record(stream)
        step()
        e1.record(stream)
    barrier()
    launches = L.ffb_launch_count(ctx) - launches0
    k_ms, k_n = ctypes.c_double(), ctypes.c_int64()
    _lib.check(ctx, L.ffb_kernel_timing_read(ctx, ctypes.byref(k_ms), ctypes.byref(k_n), 1))
    _lib.check(ctx, L.ffb_kernel_timing_enable(ctx, 0))
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in events)

    # ---- end to end through the public API ----------------------------------------------------------
    _lib.check(ctx, L.ffb_set_stream(ctx, None, 0))
    pulse = ff.PulseSequence(
        [[op, c, i] for op, c, i in zip(wl.c_opers, wl.c_coeffs, wl.c_ids)],
        [[op, c, i] for op, c, i in zip(wl.n_opers, wl.n_coeffs, wl.n_ids)],
        wl.dt, ff.Basis(wl.basis, btype='Pauli'))

    # N > 1: the user-facing call takes the GLOBAL grid; distributed.infidelity shards it (this
    # rank's block is exactly wl.omega above) and all-reduces the partial integrals.
    wl_global = base.with_omega(omega_global) if world > 1 else wl

    def e2e_step():
        pulse.cleanup('all')
        return ffd.infidelity(pulse, wl_global.spectrum, wl_global.omega) if world > 1 else \
            ff.infidelity(pulse, wl.spectrum, wl.omega)

    for _ in range(3):
        infid_e2e = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        infid_e2e = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join(timeout=1.0)

    d, n_cops, n_nops, n_basis = wl.d, len(wl.c_opers), len(wl.n_opers), len(wl.basis)
    # inputs of the fused pipeline (ffb_pulse_filter_function): operators, coefficients, dt, t, basis,
    # omega, spectrum -- one packed upload per step
    h2d = (16*(n_cops + n_nops + n_basis)*d*d + 8*(n_cops + n_nops)*G + 8*G + 8*(G + 1)
           + 8*n_omega_local + wl.spectrum.nbytes)
    d2h = (8*G*d + 16*G*d*d + 16*(G + 1)*d*d + 16*n_nops*n_basis*n_omega_local
           + 16*n_nops*n_nops*n_omega_local + 16*n_basis*n_basis + 16*n_omega_local + 8*n_nops)

    # ---- max over ranks ------------------------------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s*1e3, k_ms.value], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, k_ms_max = times.tolist()

    units_total = G*(n_per*world)        # seg*omega pairs the whole job owns per step (no halo)
    value = units_total*args.steps/(dev_ms_max*1e-3)
    e2e_value = units_total*args.steps/(e2e_ms_max*1e-3)

    # ---- parity of what was timed (cheap): infidelity of device path == API path ---------------------
    dev_inf = (total if world > 1 else dev.infidelity).cpu().numpy()
    dev_inf = dev_inf[np.argsort(wl.n_ids)]    # PulseSequence sorts operators by identifier
    parity = float(np.abs(dev_inf - np.asarray(infid_e2e).ravel()).max()
                   / np.abs(np.asarray(infid_e2e)).max())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    dfma, dmma = ctypes.c_double(), ctypes.c_double()
    _lib.check(ctx, L.ffb_measure_fp64_peak(ctx, ctypes.byref(dfma), ctypes.byref(dmma)))
    peak = max(dfma.value, dmma.value)
    kernel_ms = k_ms_max/max(1, k_n.value)
    W = wl.flops_per_seg_omega
    achieved = W*G*n_omega_local/(kernel_ms*1e-3)*1e-12
    # executed multiply-add flops of the Hermitian-pair formulation (csrc/ffb_ctrlmat.cu header): the
    # library runs the DFMA variant (thread per frequency) for <= 16 rows and d <= 3, else DMMA tiles
    rows = n_nops*n_basis
    use_dfma = rows <= 16 and 2 <= d <= 3 and os.environ.get('FFB_CTRLMAT_DFMA', '1') != '0'
    n_pairs = d*(d - 1)//2
    if use_dfma:
        rows_pad = -(-rows//4)*4
        kernel = 'ctrlmat_dfma_kernel'
        pipe = ('fp64 DFMA (thread per frequency; 64 lanes/clk/SM, the pipe DMMA shares); the '
                'operand generator adds ~52 FP64 instructions per seg*omega that are not counted in '
                'executed_tflops')
        # rows of an identity basis element carry no level-pair terms (csrc/ffb_ctrlmat.cu)
        ident0 = bool((wl.basis[0] == wl.basis[0][0, 0].real*np.eye(d)).all())
        split = (ident0 and os.environ.get('FFB_DFMA_SPLIT_IDENTITY', '1') != '0'
                 and ((d == 2 and n_basis == 4 and 2 <= n_nops <= 4)
                      or (d == 3 and n_basis == 9 and n_nops == 1)))
        pair_rows = n_nops*(n_basis - 1) if split else rows_pad
    else:
        rows_pad = -(-rows//8)*8
        tiles = rows_pad//8
        n_rb = -(-tiles//12)
        mt = next(a for a in (1, 2, 3, 4, 6, 8, 12) if a >= -(-tiles//n_rb))
        static = (d == 4 and mt in (6, 8, 12) and G >= 4
                  and os.environ.get('FFB_CTRLMAT_STATIC', '1') != '0')
        kernel = 'ctrlmat_static_kernel' if static else 'ctrlmat_main_kernel'
        rows_pad = n_rb*mt*8
        pipe = 'fp64 (DMMA.8x8x4; shares the 64 lane/clk/SM FP64 pipe with DFMA)'
        pair_rows = rows_pad
    # real multiply-adds per seg*omega: 2 per row for the diagonal unit, 4 per row and level pair
    executed = (2*rows_pad + 4*n_pairs*pair_rows)*2*G*n_omega_local/(kernel_ms*1e-3)*1e-12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
    roofline = {
        'bound': 'tensor', 'pipe': pipe,
        'kernel': kernel, 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
        'frac': achieved/peak, 'traffic': traffic,
        'peak_source': 'measured live by ffb_measure_fp64_peak (DFMA %.2f, DMMA %.2f TFLOP/s); '
                       'MEASURED_PEAKS.json has no fp64 entry' % (dfma.value, dmma.value),
        'algorithmic_flops_per_unit': W, 'units_per_launch': G*n_omega_local,
        'kernel_ms': kernel_ms, 'kernel_share_of_step': k_ms_max/dev_ms_max,
        'executed_tflops': executed, 'executed_frac': executed/peak,
        'note': 'achieved counts the reference formulation (W = 8 n_nops n_basis d^2 + 12 d^2 per '
                'seg*omega, SURVEY 8d); the kernel executes fewer flops by pairing (m,n)/(n,m) terms '
                'of Hermitian operators, so frac can exceed executed_frac (and 1.0)',
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, sample, cores = cpu_sample(base, args.cpu_seconds)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': dev_ms_max/args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(base, world, n_per),
        'clocks': sampler.result(),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_ms_max/args.steps,
                'api': 'ff.infidelity(PulseSequence, spectrum, omega) on a cold cache'},
        'gpu_launches': int(launches),
        'roofline': roofline,
        'cpu_baseline': cpu,
        'parity_device_vs_api': parity,
        'infidelity': np.asarray(infid_e2e).ravel().tolist()[:6],
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
