#!/bin/bash
TAG=${1:-r3a}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt; tail -4 gpurun_out/${TAG}_pytest.txt
for pdl in 1 0; do
  VP3D_PDL=$pdl timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-modes --train-steps 30 > gpurun_out/${TAG}_bench_pdl${pdl}.json 2> gpurun_out/${TAG}_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_pdl${pdl}.json').read().strip().splitlines()[-1])
t=d.get('train') or {}
print('PDL=${pdl}: eval',round(d['ms_per_step'],4),'e2e',round(d['e2e']['ms_per_step'],3),'train',t.get('ms_per_step'),t.get('ms_per_step_wall_incl_loss_item'),t.get('error'))
"
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_train.csv python tools/profile_steps.py train bf16 > gpurun_out/${TAG}_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches_train.csv 2>/dev/null | grep -A14 "totals"
