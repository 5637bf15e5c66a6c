#!/bin/bash
# GPU session: pack micro-benchmark, fine-grained epilogue timeline.
TAG=${1:-r2k}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 120 videopose3d_b200/_lib/dbg/pack_bench > gpurun_out/${TAG}_pack_bench.txt 2>&1
cat gpurun_out/${TAG}_pack_bench.txt
timeout 300 python tools/timeline.py fp16 0 > gpurun_out/${TAG}_timeline.txt 2>&1
grep -A3 "rep 1" gpurun_out/${TAG}_timeline.txt | cut -c1-400
