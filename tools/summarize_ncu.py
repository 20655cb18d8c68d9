#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch) into the text files kept under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_d4_v3.ncu-rep > profiles/r01_ncu_ctrlmat_d4.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__ops_path_tensor_src_fp64.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f'# {path}')
    print(f'kernel: {vals[hdr.index("Kernel Name")]}')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'{k:90s} {vals[i]:>18s} {units[i]}')
    stalls = sorted(((float(vals[i]), h) for i, h in enumerate(hdr)
                     if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')), reverse=True)
    print('warp stall reasons (warps per issue-active cycle):')
    for v, h in stalls[:8]:
        print(f'   {h.split("issue_stalled_")[1].split("_per_issue")[0]:28s} {v:.3f}')
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 2:
        hdr = rows[1]
        isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
        agg = defaultdict(lambda: [0, 0])
        for r in rows[2:]:
            toks = r[isrc].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
            op = op.split('.')[0]
            agg[op][0] += int(r[isamp] or 0)
            agg[op][1] += int(r[iex] or 0)
        tot = sum(v[0] for v in agg.values()) or 1
        print('SASS opcode mix (warp-level instructions executed, share of stall samples):')
        for op, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
            print(f'   {op:10s} executed {e:14d}   samples {s/tot:6.1%}')


if __name__ == '__main__':
    main(sys.argv[1])
