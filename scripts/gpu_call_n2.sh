#!/bin/bash
# 2-GPU pass: peer-exchange tests, the DLRM N=2 line with / without TZK_PEER_ACCUM_SIDE, kernel timelines of the sharded step.
#   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_call_n2.sh'
set -u
cd "$(dirname "$0")/.."
o=gpurun_out/n2
mkdir -p $o
(timeout 300 python -m pytest tests/test_distributed_gpu.py tests/test_peer_gpu.py -x -q -m gpu) > $o/pytest.txt 2>&1
tail -3 $o/pytest.txt
(TZK_PEER_ACCUM_SIDE=1 timeout 300 python -m pytest tests/test_distributed_gpu.py tests/test_peer_gpu.py -x -q -m gpu) > $o/pytest_accum_side.txt 2>&1
tail -3 $o/pytest_accum_side.txt
port=29711
for side in 0 1; do
  port=$((port + 1))
  TZK_PEER_ACCUM_SIDE=$side timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --trace $o/trace_side$side.json \
      > $o/bench_dlrm_side$side.json 2> $o/bench_dlrm_side$side.err
  python - $o/bench_dlrm_side$side.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e6, 2), "M  verify", d.get("verify"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  python scripts/trace_summary.py $o/trace_side$side.json > $o/trace_side$side.txt 2>&1
  rm -f $o/trace_side$side.json
done
head -100 $o/trace_side0.txt
