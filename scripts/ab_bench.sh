#!/bin/bash
# A/B of environment switches on the headline step, one bench.py run per setting, results in gpurun_out/ab/.
#
#   gpurun --timeout 900 -- 'bash scripts/ab_bench.sh "" TZK_SMALL_LINEAR_BWD=1 TZK_GEMM3X=1 "TZK_GEMM3X=1 TZK_GEMM3X_STACK=1" "TZK_GEMM3X=1 TZK_GEMM3X_STACK=1 TZK_GEMM3X_TW=8" "TZK_GEMM3X=1 TZK_GEMM3X_STACK=1 TZK_GEMM3X_SPLIT=1" "TZK_GEMM3X=1 TZK_GEMM3X_STACK=1 TZK_GEMM3X_SPLIT=1 TZK_GEMM3X_RAW=1 TZK_GEMM3X_PREFETCH=1" "TZK_GEMM3X=1 TZK_GEMM3X_RING=1 TZK_GEMM3X_PREFETCH=1" TZK_L2_PERSIST=1'
#
# Each argument is a (possibly empty) space-separated list of VAR=value; "" is the baseline.  Prints one line per
# setting: ms/step (inputs resident), e2e ms/step, fused_bwd apply µs.  Extras / CPU baseline / Zipf are skipped.
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/ab
mkdir -p "$out"
i=0
for setting in "$@"; do
  i=$((i + 1))
  f="$out/ab_$i.json"
  env $setting python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-zipf --no-extras >"$f" 2>"$out/ab_$i.err"
  python - "$f" "$setting" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = (d.get("roofline") or {}).get("kernels", {})
    fb = k.get("fused_bwd", {})
    print(f"{sys.argv[2] or '(baseline)':<48} {d['ms_per_step']:.4f} ms/step   e2e {d['e2e']['ms_per_step']:.4f}   "
          f"fused_bwd apply {1e3 * fb.get('apply_ms', float('nan')):.1f} us")
except Exception as e:
    print(f"{sys.argv[2] or '(baseline)':<48} FAILED: {e}")
PY
done
