#!/bin/bash
# One 1-GPU validation + measurement pass (gpurun runs whatever this file says at snapshot time):
#   gpurun --timeout 1700 -- 'bash scripts/gpu_call_n1.sh'
# TZK_EXPERIMENTAL=1 switches on every path that has not been validated on hardware yet (interleaved arenas, dW tile kernel,
# staged scan stores, tensor-core interaction) and un-skips their tests.
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/c3
mkdir -p $o
nvidia-smi -L > $o/gpus.txt
(time TZK_EXPERIMENTAL=1 python -m pytest tests -x -q -m gpu) > $o/pytest_experimental.txt 2>&1
tail -4 $o/pytest_experimental.txt
(time python -m pytest tests -x -q -m gpu) > $o/pytest_default.txt 2>&1
tail -4 $o/pytest_default.txt
bash scripts/ab_bench.sh "" TZK_EXPERIMENTAL=1 "TZK_EXPERIMENTAL=1 TZK_INTERACT_TC=0" "TZK_EXPERIMENTAL=1 TZK_INTERLEAVE=0" \
    "TZK_EXPERIMENTAL=1 TZK_SMALL_LINEAR_DW=0" "TZK_EXPERIMENTAL=1 TZK_SCAN_STAGED=0" > $o/ab.txt 2>&1
cat $o/ab.txt
for tc in 0 1; do
  TZK_EXPERIMENTAL=1 TZK_INTERACT_TC=$tc ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $o/launches_tc$tc.csv python bench.py --steps 3 --warmup 3 --ring 2 --no-cpu-baseline --no-zipf --no-extras \
      > $o/ncu_bench_tc$tc.log 2>&1
  python scripts/summarize_launches.py $o/launches_tc$tc.csv > $o/launch_summary_tc$tc.txt 2>&1
done
head -8 $o/launch_summary_tc1.txt
grep -E "small_linear|reduce2|fused_apply|pooled_gather|scan_tiles|dot_interact" $o/launch_summary_tc0.txt $o/launch_summary_tc1.txt | cut -c1-150
TZK_EXPERIMENTAL=1 timeout 400 python scripts/ncu_traffic.py r2b > $o/ncu_traffic.txt 2>&1
tail -12 $o/ncu_traffic.txt
TZK_EXPERIMENTAL=1 python bench.py --steps 20 --warmup 5 > $o/bench_experimental.json 2> $o/bench_experimental.err
tail -c 600 $o/bench_experimental.json
