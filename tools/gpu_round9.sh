#!/bin/bash
# GPU session: pack micro-benchmark, op/parity tests, quick bench, full ncu capture of one eval forward.
TAG=${1:-r2l}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 120 videopose3d_b200/_lib/dbg/pack_bench > gpurun_out/${TAG}_pack_bench.txt 2>&1
cat gpurun_out/${TAG}_pack_bench.txt
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/${TAG}_ops.txt 2>&1
echo "ops+parity exit $?"; tail -4 gpurun_out/${TAG}_ops.txt
timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-train --no-modes > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('bench value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),'dom frac',round(d['roofline']['frac'],3),'step frac',round(d['roofline_step']['frac'],3))
"
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/${TAG}_full_eval_fp16 -f python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
