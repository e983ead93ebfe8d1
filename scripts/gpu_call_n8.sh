#!/bin/bash
# N-GPU pass (default 8): the DLRM row-wise line with / without the split gather, then the other BASELINE configs.
#   gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_call_n8.sh 8'
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
o=gpurun_out/n$N
mkdir -p $o
nvidia-smi -L > $o/gpus.txt
port=29800
line() {  # name, env..., -- args...
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  port=$((port + 1))
  env "${envs[@]}" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus "$N" "$@" > $o/$name.json 2> $o/$name.err
  python - $o/$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e6, 2), "M  e2e",
          round(d["e2e"]["value"] / 1e6, 2), "M  verify", d.get("verify"), " gather ms", r.get("ms"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
line dlrm TZK_NOOP=1 -- --steps 20 --warmup 5
line dlrm_split TZK_PEER_SPLIT_GATHER=1 TZK_PEER_MIRROR_CHUNKED=1 -- --steps 20 --warmup 5 --no-verify
line deepfm TZK_NOOP=1 -- --model deepfm_criteo --sharding table_wise --batch-size 32768 --steps 20 --warmup 5
line mmoe TZK_NOOP=1 -- --model mmoe_taobao --sharding mixed --batch-size 8192 --steps 20 --warmup 5
line din TZK_NOOP=1 -- --model multi_tower_din_taobao --sharding mixed --batch-size 8192 --steps 20 --warmup 5
