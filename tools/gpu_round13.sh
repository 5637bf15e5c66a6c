#!/bin/bash
TAG=${1:-r2u}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_train.csv python tools/profile_steps.py train bf16 > gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
python tools/summarize_launches.py gpurun_out/${TAG}_launches_train.csv 2>/dev/null | tail -32
timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-modes > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('bench value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'step frac',round(d['roofline_step']['frac'],3))
t=d.get('train') or {}
print('  train',t.get('ms_per_step'),t.get('ms_per_step_wall_incl_loss_item'),t.get('error'))
"
