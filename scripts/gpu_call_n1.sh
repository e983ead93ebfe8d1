#!/bin/bash
# One 1-GPU validation + measurement pass (gpurun runs whatever this file says at snapshot time):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_call_n1.sh'
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/c6
mkdir -p $o
(time TZK_EXPERIMENTAL=1 python -m pytest tests -x -q -m gpu) > $o/pytest_experimental.txt 2>&1
tail -4 $o/pytest_experimental.txt
bash scripts/ab_bench.sh "" TZK_BWD_HEADS=1 > $o/ab.txt 2>&1
cat $o/ab.txt
TZK_BWD_HEADS=1 timeout 400 python scripts/ncu_traffic.py r2e > $o/ncu_traffic.txt 2>&1
tail -4 $o/ncu_traffic.txt | cut -c1-200
