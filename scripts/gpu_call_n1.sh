#!/bin/bash
# Last short 1-GPU pass of the round (3 GPU-minutes left): default-path tests first, then the headline line, then the fused
# tail A/B if the box is still there.
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/c8
mkdir -p $o
(time timeout 100 python -m pytest tests -x -q -m gpu) > $o/pytest_default.txt 2>&1
tail -3 $o/pytest_default.txt
timeout 60 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-zipf --no-extras > $o/bench_default.json 2> $o/bench_default.err
tail -c 300 $o/bench_default.json
TZK_FUSED_TAIL=1 timeout 60 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-zipf --no-extras > $o/bench_tail.json 2> $o/bench_tail.err
tail -c 300 $o/bench_tail.json
(TZK_EXPERIMENTAL=1 timeout 60 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -m gpu -k "tower_tail or train_steps or graph") > $o/pytest_tail.txt 2>&1
tail -3 $o/pytest_tail.txt
