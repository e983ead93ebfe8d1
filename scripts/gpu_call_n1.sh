#!/bin/bash
# One 1-GPU validation + measurement pass (gpurun runs whatever this file says at snapshot time):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_call_n1.sh'
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/c7
mkdir -p $o
(time TZK_EXPERIMENTAL=1 python -m pytest tests -x -q -m gpu) > $o/pytest_experimental.txt 2>&1
tail -4 $o/pytest_experimental.txt
(time python -m pytest tests -x -q -m gpu) > $o/pytest_default.txt 2>&1
tail -4 $o/pytest_default.txt
bash scripts/ab_bench.sh "" TZK_FUSED_TAIL=1 TZK_DLRM_BOTTOM_STREAM=1 "TZK_FUSED_TAIL=1 TZK_DLRM_BOTTOM_STREAM=1" > $o/ab.txt 2>&1
cat $o/ab.txt
TZK_FUSED_TAIL=1 TZK_DLRM_BOTTOM_STREAM=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches.csv \
    python bench.py --steps 3 --warmup 3 --ring 2 --no-cpu-baseline --no-zipf --no-extras > $o/ncu_bench.log 2>&1
python scripts/summarize_launches.py $o/launches.csv > $o/launch_summary.txt 2>&1
head -60 $o/launch_summary.txt | cut -c1-120
