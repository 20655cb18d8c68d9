"""Run the REFERENCE's own test-suite against filter_functions_b200 (GPU box).

The unmodified tests of qutech/filter_functions (staged byte-identically under ``baseline/_ref_tests`` by
``baseline/install_reference.py``; git-ignored, never part of this repository's history) are collected
with ``import filter_functions`` resolving to this package (``tools/ref_alias``).  They are mostly
unseeded randomised property tests, so every run exercises new inputs.  Tests of components that are
out of scope here (gradients, plotting, extend / remap, second-order filter functions, parts of
``util`` / ``basis``) fail or error by construction; the summary lists, per test file, what passed, and
for every test that did not, the first line of the reason -- so that an in-scope failure cannot hide.

    python tools/run_reference_tests.py [--files test_precision.py,test_sequencing.py] [--out summary.json]
"""
import argparse
import json
import os
import subprocess
import sys
import xml.etree.ElementTree as ET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, 'baseline', '_ref_tests')
# reasons that mark a test as exercising a component outside SURVEY.md section 8
OUT_OF_SCOPE = [
    ('second-order filter functions / frequency shifts (SURVEY 2, row 11)',
     ('Second-order', 'second order', 'calculate_frequency_shifts', 'second_order')),
    ('extend / remap and the tensor-chain helpers they use (SURVEY 2, row 14)',
     ("no attribute 'extend'", "no attribute 'remap'", 'tensor_insert', 'tensor_merge', 'tensor_transpose')),
    ('Hilbert-space noise-operator variants (SURVEY 2, row 12)', ('calculate_noise_operators',)),
    ("private helpers of the reference's NumPy implementation (their work happens inside the kernels here)",
     ("'_get_integrand'", "'_first_order_integral'", "'_transform_hamiltonian'", "'_second_order_integral'")),
    ('optional dependency absent from the image', ('qutip', 'matplotlib')),
]


def classify(reason):
    for label, needles in OUT_OF_SCOPE:
        if any(n in reason for n in needles):
            return label
    return None


DEFAULT = ['test_precision.py', 'test_sequencing.py', 'test_core.py', 'test_superoperator.py',
           'test_util.py', 'test_basis.py']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--files', default=','.join(DEFAULT))
    ap.add_argument('--out', default=None)
    ap.add_argument('--timeout', type=int, default=900)
    ap.add_argument('-k', default=None)
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(STAGED, 'tests')):
        print(json.dumps({'unavailable': 'reference tests not staged (baseline/install_reference.py)'}))
        return
    env = dict(os.environ)
    env['PYTHONPATH'] = os.pathsep.join([STAGED, os.path.join(ROOT, 'tools', 'ref_alias'), ROOT,
                                         os.path.join(ROOT, 'oracle', 'shim'), env.get('PYTHONPATH', '')])
    summary = {}
    for f in [x for x in args.files.split(',') if x]:
        xml = os.path.join('/tmp', f'ref_{f}.xml')
        cmd = [sys.executable, '-m', 'pytest', os.path.join('tests', f), '-q', '-x' if False else '-q',
               '-p', 'no:cacheprovider', '-o', 'addopts=', '--timeout', '300', f'--junitxml={xml}',
               '--rootdir', STAGED, '-W', 'ignore']
        if args.k:
            cmd += ['-k', args.k]
        try:
            res = subprocess.run(cmd, cwd=STAGED, env=env, capture_output=True, text=True,
                                 timeout=args.timeout)
            tail = (res.stdout or '').strip().splitlines()[-1:] + (res.stderr or '').strip().splitlines()[-2:]
        except subprocess.TimeoutExpired:
            summary[f] = {'timeout': args.timeout}
            continue
        entry = {'passed': 0, 'failed': 0, 'errors': 0, 'skipped': 0, 'in_scope_failures': 0,
                 'not_passed': [], 'pytest': tail}
        if os.path.exists(xml):
            for case in ET.parse(xml).getroot().iter('testcase'):
                name = f"{case.get('classname', '').split('.')[-1]}::{case.get('name')}"
                bad = None
                for kind in ('failure', 'error', 'skipped'):
                    node = case.find(kind)
                    if node is not None:
                        bad = (kind, (node.get('message') or node.text or '').strip().splitlines()[:1])
                        break
                if bad is None:
                    entry['passed'] += 1
                else:
                    key = {'failure': 'failed', 'error': 'errors', 'skipped': 'skipped'}[bad[0]]
                    entry[key] += 1
                    reason = (bad[1][0] if bad[1] else '')[:200]
                    scope = classify(reason)
                    entry['not_passed'].append({'test': name, 'outcome': bad[0], 'reason': reason,
                                                'out_of_scope': scope})
                    if scope is None and bad[0] != 'skipped':
                        entry['in_scope_failures'] = entry.get('in_scope_failures', 0) + 1
            os.remove(xml)
        summary[f] = entry
    text = json.dumps(summary, indent=1)
    if args.out:
        with open(args.out, 'w') as fh:
            fh.write(text + '\n')
    for f, e in summary.items():
        print(f, {k: v for k, v in e.items() if k not in ('not_passed',)})
        for n in e.get('not_passed', []):
            print('   ', n['outcome'], n['test'], '--', n['reason'],
                  '[out of scope: %s]' % n['out_of_scope'] if n['out_of_scope'] else '[IN SCOPE]')
    tot = {k: sum(e.get(k, 0) for e in summary.values()) for k in ('passed', 'failed', 'errors', 'skipped',
                                                                   'in_scope_failures')}
    print('TOTAL', tot)


if __name__ == '__main__':
    main()
