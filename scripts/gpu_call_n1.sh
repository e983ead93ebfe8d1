#!/bin/bash
# One 1-GPU validation + measurement pass (gpurun runs whatever this file says at snapshot time):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_call_n1.sh'
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/c3
mkdir -p $o
nvidia-smi -L > $o/gpus.txt
(time python -m pytest tests -x -q -m gpu) > $o/pytest.txt 2>&1
tail -4 $o/pytest.txt
bash scripts/ab_bench.sh "" TZK_INTERACT_TC=1 TZK_INTERLEAVE=0 TZK_SMALL_LINEAR_DW=0 "TZK_INTERACT_TC=1 TZK_INTERLEAVE=0" > $o/ab.txt 2>&1
cat $o/ab.txt
for tc in 0 1; do
  TZK_INTERACT_TC=$tc ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/launches_tc$tc.csv \
      python bench.py --steps 3 --warmup 3 --ring 2 --no-cpu-baseline --no-zipf --no-extras > $o/ncu_bench_tc$tc.log 2>&1
  python scripts/summarize_launches.py $o/launches_tc$tc.csv > $o/launch_summary_tc$tc.txt 2>&1
done
head -8 $o/launch_summary_tc0.txt
grep -E "small_linear|reduce2|fused_apply|pooled_gather|scan_tiles|dot_interact" $o/launch_summary_tc0.txt $o/launch_summary_tc1.txt | cut -c1-150
timeout 400 python scripts/ncu_traffic.py r2b > $o/ncu_traffic.txt 2>&1
tail -12 $o/ncu_traffic.txt
