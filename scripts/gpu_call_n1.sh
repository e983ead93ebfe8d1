#!/bin/bash
# ncu traffic of the gather + fused backward on the final sources (-> profiles/ncu_traffic.json, read by bench.py)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/c9
timeout 90 python scripts/ncu_traffic.py r2f > gpurun_out/c9/ncu_traffic.txt 2>&1
tail -14 gpurun_out/c9/ncu_traffic.txt | cut -c1-160
