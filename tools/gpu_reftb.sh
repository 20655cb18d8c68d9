#!/bin/bash
# tracebacks of selected reference tests against this package
cd baseline/_ref_tests
export PYTHONPATH=$GRAFT_REPO_ROOT/baseline/_ref_tests:$GRAFT_REPO_ROOT/tools/ref_alias:$GRAFT_REPO_ROOT:$GRAFT_REPO_ROOT/oracle/shim
timeout 900 python -m pytest tests -q -p no:cacheprovider -o addopts= --rootdir . -W ignore --tb=short -k "$1" 2>&1 | grep -v "^$" | tail -${2:-200}
