#!/bin/bash
# tile geometry sweep: prints value + per-stage ms
for cfg in "48 32" "32 32" "64 64" "96 64" "96 128" "128 128" "160 128"; do
  set -- $cfg
  out=$(TFR_TILE_KB=$1 TFR_TILE_THREADS=$2 python bench.py --steps 6 --warmup 3 --batch-mib 256 --no-e2e --no-cpu 2>&1 | tail -1)
  echo "$1KB x $2thr: $(echo "$out" | python -c "import sys,json; l=json.loads(sys.stdin.read()); print(round(l['value'],1),'GB/s', 'ms/step',round(l['ms_per_step'],3), l['step_hbm']['stage_ms_per_step'], 'launches', l['gpu_launches'])" 2>&1 | tail -1)"
done
