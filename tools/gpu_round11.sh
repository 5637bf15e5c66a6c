#!/bin/bash
# GPU session: the whole GPU suite, smoke, bench (incl. pair-128 off for comparison), launch list, timeline.
TAG=${1:-r2s}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt
tail -6 gpurun_out/${TAG}_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-modes > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
for f in ['bench']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),'dom frac',round(d['roofline']['frac'],3),'step frac',round(d['roofline_step']['frac'],3))
        t=d.get('train') or {}
        if t: print('  train',t.get('ms_per_step'),t.get('ms_per_step_wall_incl_loss_item'),t.get('error'))
    except Exception as e: print(f,'ERR',e)
"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval.csv python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches_eval.csv 2>/dev/null | head -12
timeout 300 python tools/timeline.py fp16 > gpurun_out/${TAG}_timeline.txt 2>&1
grep "^# rep" gpurun_out/${TAG}_timeline.txt
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
