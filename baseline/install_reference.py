"""Stage the UNMODIFIED reference (qutech/filter_functions, pure Python) under ``baseline/_ref`` so
that ``bench.py --impl reference`` can time the reference's own CPU implementation on the GPU box,
where ``/root/reference`` does not exist.  ``baseline/_ref`` is git-ignored (never part of the
history) but travels with ``gpurun``.

1. ``pip install --no-index --no-build-isolation --target baseline/_ref /root/reference`` is tried
   first.  In this image it fails: the reference's build backend (``hatchling``) is not installed and
   there is no network.
2. Fallback: the package is pure Python with no generated files, so installing it IS copying its
   package directory; ``filter_functions/`` is copied verbatim (byte-identical, checked).

The two run-time dependencies that are missing from the image (``opt_einsum``, ``sparse``; unpinned
in the reference's ``pyproject.toml:28-34``) are provided by the NumPy stand-ins in ``oracle/shim``
(SURVEY.md section 8c); they are put on ``sys.path`` by the caller, not copied here.
"""
import filecmp
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, '_ref')
SOURCE = '/root/reference'


def install(verbose: bool = True) -> str:
    """Returns 'pip', 'copy', 'present' or 'unavailable'."""
    pkg = os.path.join(TARGET, 'filter_functions')
    if not os.path.isdir(SOURCE):
        return 'present' if os.path.isdir(pkg) else 'unavailable'
    src_pkg = os.path.join(SOURCE, 'filter_functions')
    if os.path.isdir(pkg) and not filecmp.dircmp(src_pkg, pkg, ignore=['__pycache__']).diff_files:
        return 'present'
    os.makedirs(TARGET, exist_ok=True)
    cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps',
           '--find-links', '/opt/wheelhouse', '--target', TARGET, SOURCE]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode == 0 and os.path.isdir(pkg):
        return 'pip'
    if verbose:
        tail = (res.stderr or res.stdout).strip().splitlines()[-1:]
        print(f'[baseline] pip install failed ({tail}); copying the pure-Python package', file=sys.stderr)
    shutil.rmtree(pkg, ignore_errors=True)
    shutil.copytree(src_pkg, pkg, ignore=shutil.ignore_patterns('__pycache__'))
    cmp = filecmp.dircmp(src_pkg, pkg, ignore=['__pycache__'])
    assert not cmp.diff_files and not cmp.left_only, (cmp.diff_files, cmp.left_only)
    return 'copy'


def install_tests(verbose: bool = True) -> str:
    """Stage the reference's own test-suite (``tests/`` + the ``examples/data`` fixtures it loads) under
    ``baseline/_ref_tests`` (git-ignored, travels with ``gpurun``), byte-identical, so that
    ``tools/run_reference_tests.py`` can run it against THIS package on the GPU box.  Returns 'copy',
    'present' or 'unavailable'."""
    target = os.path.join(HERE, '_ref_tests')
    if not os.path.isdir(SOURCE):
        return 'present' if os.path.isdir(os.path.join(target, 'tests')) else 'unavailable'
    state = 'present'
    for sub in ('tests', os.path.join('examples', 'data')):
        src, dst = os.path.join(SOURCE, sub), os.path.join(target, sub)
        if os.path.isdir(dst) and not filecmp.dircmp(src, dst, ignore=['__pycache__']).diff_files \
                and not filecmp.dircmp(src, dst, ignore=['__pycache__']).left_only:
            continue
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__'))
        state = 'copy'
    return state


if __name__ == '__main__':
    print(install())
    print('tests:', install_tests())
