import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 GPU (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def engine():
    """Builds (if needed) and loads libffb200; GPU tests call through it."""
    import __graft_entry__ as entry
    entry.build()
    import filter_functions_b200 as ff
    return ff
