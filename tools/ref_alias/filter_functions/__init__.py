"""``import filter_functions`` -> filter_functions_b200, for running the REFERENCE's own test-suite
against this package (tools/run_reference_tests.py).  Test infrastructure only: nothing in the product
imports it.  Modules this package does not build (``analytic``: closed-form filter functions the
reference's tests use as known answers; ``types``) are taken from the staged, unmodified reference under
``baseline/_ref`` -- as test oracles, never on the product path.  Everything else that is out of scope
(gradient, plotting, extend / remap, second order) is simply absent, so the tests that need it fail and
are reported as such."""
import importlib.util
import os
import sys

import filter_functions_b200 as _ff
from filter_functions_b200 import *  # noqa: F401,F403
from filter_functions_b200 import basis, numeric, pulse_sequence, superoperator, util  # noqa: F401

__version__ = _ff.__version__
for _name in ('basis', 'numeric', 'pulse_sequence', 'superoperator', 'util'):
    sys.modules[f'{__name__}.{_name}'] = getattr(_ff, _name)

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', '..', 'baseline', '_ref',
                    'filter_functions')
for _name in ('analytic', 'types'):
    _path = os.path.join(_REF, _name + '.py')
    if os.path.exists(_path):
        _spec = importlib.util.spec_from_file_location(f'{__name__}.{_name}', _path)
        _mod = importlib.util.module_from_spec(_spec)
        sys.modules[f'{__name__}.{_name}'] = _mod
        _spec.loader.exec_module(_mod)
        globals()[_name] = _mod
