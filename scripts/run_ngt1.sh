#!/bin/bash
# One N>1 bench line per BASELINE config (torchrun, one rank per GPU), results in gpurun_out/ngt1/.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/run_ngt1.sh 2'
#   gpurun --gpus 8 --timeout 600 -- 'bash scripts/run_ngt1.sh 8 dlrm'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
only=${2:-all}
out=gpurun_out/ngt1
mkdir -p "$out"
port=29610
run() {  # name, args...
  local name=$1; shift
  port=$((port + 1))
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus "$N" "$@" >"$out/n${N}_${name}.json" 2>"$out/n${N}_${name}.err"
  echo "rc=$? $name: $(tail -c 600 "$out/n${N}_${name}.json")"
}
case "$only" in all|dlrm) run dlrm --steps 20 --warmup 5 ;; esac
case "$only" in all|deepfm) run deepfm --model deepfm_criteo --sharding table_wise --batch-size 32768 --steps 20 --warmup 5 ;; esac
case "$only" in all|mmoe) run mmoe --model mmoe_taobao --sharding mixed --batch-size 8192 --steps 20 --warmup 5 ;; esac
case "$only" in all|din) run din --model multi_tower_din_taobao --sharding mixed --batch-size 8192 --steps 20 --warmup 5 ;; esac
