"""bench.py's output contract, checked on the CPU through the reference arm (the GPU arm prints through
the same code path): exactly ONE line on stdout, JSON, with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                           '--steps', '1', '--warmup', '0', '--cpu-seconds', '1'],
                          capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = proc.stdout.splitlines()
    assert len(lines) == 1, proc.stdout[:2000]
    line = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
                'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline',
                'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['value'] > 0 and line['unit'] == 'seg*omega/s'
    # the headline workload is the north_star target shape; the arm runs the unmodified reference when it
    # is staged under baseline/_ref (always the case where /root/reference exists), else the oracle port
    assert line['config']['workload'].startswith('d4') and line['scaling'] == 'strong'
    staged = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'filter_functions'))
    assert line['cpu_baseline']['kind'] == ('reference' if staged else 'port')
    assert line['cpu_baseline']['cores'] >= 1
    assert line['extrapolated'] is True and 'EXTRAPOLATED' in line['ms_per_step_note']
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    assert line['vs_baseline'] is None and line['dtype'] == 'f64' and line['data'] == 'synthetic'
