#!/bin/bash
TAG=${1:-r2t}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for st in 0 4 3 2; do
  VP3D_DBG_STAGES=$st timeout 300 python tools/timeline.py fp16 > gpurun_out/${TAG}_timeline_st${st}.txt 2>&1
  echo "== stages $st"; grep -A100 "rep 1" gpurun_out/${TAG}_timeline_st${st}.txt | grep " us:" | grep cta0 | awk '{printf "%s ", $1; for(i=4;i<=NF;i++){split($i,a,"="); v[a[1]]=a[2]}; printf "land0 %.2f mma_t0 %.2f first_tile %.2f mma_span %.2f\n", v["land0"], v["mma_t0"], v["mma_t0"]-v["land0"], v["mma_tN"]-v["land0"]}'
done
