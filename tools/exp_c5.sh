#!/bin/bash
# C5 concatenation variants: from_atomic tile height and stack mode
for lt in 16 8; do for st in 0 1; do
  echo "LT=$lt STACK=$st"; FFB_FA_LT=$lt FFB_CONCAT_STACK=$st python tools/run_c5.py 10000 5 | tail -2
done; done
