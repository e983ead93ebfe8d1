#!/bin/bash
# One 1-GPU validation + measurement pass (gpurun runs whatever this file says at snapshot time):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_call_n1.sh'
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/c5
mkdir -p $o
(time python -m pytest tests -x -q -m gpu) > $o/pytest_default.txt 2>&1
tail -4 $o/pytest_default.txt
bash scripts/ab_bench.sh "" TZK_INTERACT_TC_BWD=0 > $o/ab.txt 2>&1
cat $o/ab.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches.csv \
    python bench.py --steps 3 --warmup 3 --ring 2 --no-cpu-baseline --no-zipf --no-extras > $o/ncu_bench.log 2>&1
python scripts/summarize_launches.py $o/launches.csv > $o/launch_summary.txt 2>&1
head -8 $o/launch_summary.txt
grep -E "fused_apply|pooled_gather|dot_interact" $o/launch_summary.txt | cut -c1-150
timeout 400 python scripts/ncu_traffic.py r2d > $o/ncu_traffic.txt 2>&1
tail -4 $o/ncu_traffic.txt | cut -c1-200
python bench.py --steps 20 --warmup 5 > $o/bench.json 2> $o/bench.err
tail -c 700 $o/bench.json
