#!/bin/bash
TAG=${1:-r2y}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_step_ops.py -m gpu -q > gpurun_out/${TAG}_train_tests.txt 2>&1; echo "tests exit $?"; tail -4 gpurun_out/${TAG}_train_tests.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cudnn --no-modes --train-steps 30 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
t=d.get('train') or {}
print('train',t.get('ms_per_step'),t.get('ms_per_step_wall_incl_loss_item'),t.get('error'))
"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_train.csv python tools/profile_steps.py train bf16 > gpurun_out/${TAG}_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches_train.csv 2>/dev/null | grep -A16 "totals"
