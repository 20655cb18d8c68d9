"""The C-ABI library: builds, loads, exports every symbol include/ffb200.h declares, and fails
loudly (no CPU fallback) when there is no GPU.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as entry
    entry.build()
    from filter_functions_b200 import _lib
    return _lib


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'ffb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ffb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib):
    handle = lib.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(handle, name), f'{name} declared in include/ffb200.h but not exported'
    # the Python binding table covers the header exactly
    assert sorted(lib.EXPORTED_SYMBOLS) == names


def test_library_is_in_tree(lib):
    assert lib.library_path.startswith(ROOT)
    assert os.path.exists(lib.library_path)


def test_version_and_error_codes(lib):
    assert b'sm_100a' in lib.lib().ffb_version()
    assert lib.FFB_OK == 0 and lib.FFB_EINVAL < 0


def test_sass_is_sm100_with_fp64_tensor_instructions(lib):
    """The shipped kernels are sm_100a SASS and the control-matrix kernel uses DMMA."""
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '-lelf', lib.library_path], capture_output=True, text=True)
    assert 'sm_100a' in out.stdout
    sass = subprocess.run([cuobjdump, '-sass', lib.library_path], capture_output=True, text=True)
    assert 'DMMA.8x8x4' in sass.stdout
    assert 'LDGSTS' in sass.stdout  # cp.async staging of the operand stream


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    handle = ctypes.c_void_p()
    rc = lib.lib().ffb_init(ctypes.byref(handle), 0)
    assert rc == lib.FFB_ENODEVICE
    assert b'no CPU fallback' in lib.lib().ffb_last_error(None)
    import filter_functions_b200 as ff
    with pytest.raises(lib.FFBError):
        ff.numeric.calculate_filter_function(np.zeros((1, 4, 3), dtype=complex))
    with pytest.raises(lib.FFBError):
        ff.numeric.diagonalize(np.zeros((2, 2, 2), dtype=complex), [1.0, 1.0])


def test_product_never_imports_oracle():
    """Nothing under filter_functions_b200/ may import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, 'filter_functions_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'ff_oracle' not in text and 'import oracle' not in text, f
                assert '/root/reference' not in text, f
