#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
tools/microbench/fp64_peaks 2>/dev/null | grep -E "dmma_lds|mixed"
echo "--- baseline"
bash tools/exp_build_run.sh "" c2 c3 d4
echo "--- NOGEN"
bash tools/exp_build_run.sh "-DFFB_EXP_NOGEN" c2 c3 d4
