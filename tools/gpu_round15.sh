#!/bin/bash
TAG=${1:-r2w}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for rows in 1000000000 0 4096 12288 30000; do
  VP3D_FUSE_BNB_ROWS=$rows timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cudnn --no-modes --train-steps 30 > gpurun_out/${TAG}_bench_${rows}.json 2> gpurun_out/${TAG}_bench.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${rows}.json').read().strip().splitlines()[-1])
t=d.get('train') or {}
print('fuse rows <= ${rows}: train',t.get('ms_per_step'),t.get('ms_per_step_wall_incl_loss_item'),t.get('error'))
"
done
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q > gpurun_out/${TAG}_train_tests.txt 2>&1; tail -3 gpurun_out/${TAG}_train_tests.txt
