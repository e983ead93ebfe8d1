#!/bin/bash
# First GPU contact of every draft in this directory, each under its own timeout (a wrong mbarrier phase or flag
# hangs a kernel), logs under gpurun_out/experimental/.
#
#   gpurun --timeout 900 -- 'bash scripts/experimental/run_all.sh'              # single-GPU drafts
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/experimental/run_all.sh peer' # + the peer-memory sparse step
set -u
cd "$(dirname "$0")/../.."
out=gpurun_out/experimental
mkdir -p "$out"
run() {  # name, timeout seconds, command...
  local name=$1 t=$2
  shift 2
  echo "== $name" | tee "$out/$name.log"
  timeout "$t" "$@" >>"$out/$name.log" 2>&1
  echo "exit $?" | tee -a "$out/$name.log"
  tail -n 6 "$out/$name.log"
}
run tower_bwd2 240 python scripts/experimental/try_tower_bwd2.py 65536
run gemm3x_fwd_small 120 python scripts/experimental/try_gemm3x.py 256 fwd
run gemm3x_fwd 120 python scripts/experimental/try_gemm3x.py 65536 fwd
run gemm3x_dgrad 120 python scripts/experimental/try_gemm3x.py 65536 dgrad
run gemm3x_wgrad 120 python scripts/experimental/try_gemm3x.py 65536 wgrad
if [ "${1:-}" = "peer" ]; then
  run peer 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29611 scripts/experimental/try_peer.py 8192
fi
nvidia-smi --query-gpu=name,clocks.sm,clocks_throttle_reasons.active --format=csv >"$out/nvidia_smi.csv" 2>&1
exit 0
