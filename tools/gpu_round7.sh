#!/bin/bash
# GPU session: op + parity tests of the new pack / lean epilogue, quick bench, launch list, timelines.
TAG=${1:-r2j}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/${TAG}_ops.txt 2>&1
echo "ops+parity exit $?"; tail -4 gpurun_out/${TAG}_ops.txt
timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-train --no-modes > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval.csv python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_prof.log 2>&1
for exp in 0 1 3; do
  timeout 300 python tools/timeline.py fp16 $exp > gpurun_out/${TAG}_timeline_exp${exp}.txt 2>&1
done
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('bench value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),'dom frac',round(d['roofline']['frac'],3),'step frac',round(d['roofline_step']['frac'],3))
"
python tools/summarize_launches.py gpurun_out/${TAG}_launches_eval.csv 2>/dev/null | tail -16
grep "^# rep" gpurun_out/${TAG}_timeline_exp*.txt
