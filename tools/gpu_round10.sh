#!/bin/bash
# GPU session: op/parity tests (each file under its own timeout: a protocol bug shows up as a hang),
# quick bench with and without the lean epilogue, launch list, timeline.
TAG=${1:-r2m}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -m gpu -q -x > gpurun_out/${TAG}_ops.txt 2>&1
echo "ops exit $?"; tail -4 gpurun_out/${TAG}_ops.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${TAG}_parity.txt 2>&1
echo "parity exit $?"; tail -4 gpurun_out/${TAG}_parity.txt
timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-train --no-modes > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
VP3D_LEAN=0 timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-train --no-modes > gpurun_out/${TAG}_bench_nolean.json 2>> gpurun_out/${TAG}_bench.err
python -c "
import json
for f in ['bench','bench_nolean']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),'dom frac',round(d['roofline']['frac'],3),'step frac',round(d['roofline_step']['frac'],3))
    except Exception as e: print(f,'ERR',e)
"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval.csv python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches_eval.csv 2>/dev/null | head -14
timeout 300 python tools/timeline.py fp16 > gpurun_out/${TAG}_timeline.txt 2>&1
grep "^# rep" gpurun_out/${TAG}_timeline.txt
