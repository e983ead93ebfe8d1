#!/bin/bash
# 2-GPU pass: peer-exchange tests, the DLRM N=2 line, a kernel timeline of the sharded step.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_call_n2.sh'
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/n2
mkdir -p $o
(timeout 400 python -m pytest tests/test_distributed_gpu.py tests/test_peer_gpu.py -x -q -m gpu) > $o/pytest.txt 2>&1
tail -3 $o/pytest.txt
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
    bench.py --gpus 2 --steps 20 --warmup 5 --trace $o/trace_dlrm.json > $o/bench_dlrm.json 2> $o/bench_dlrm.err
tail -c 400 $o/bench_dlrm.json
python scripts/trace_summary.py $o/trace_dlrm.json > $o/trace_dlrm.txt 2>&1
head -90 $o/trace_dlrm.txt
rm -f $o/trace_dlrm.json
