#!/bin/bash
# GPU session: remaining tests (no -x), bench with train block, ncu full capture of the eval forward.
TAG=${1:-r2b}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_step_ops.py tests/test_gpu_train.py tests/test_gpu_run_py.py -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt
tail -15 gpurun_out/${TAG}_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/${TAG}_full_eval_fp16 -f python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_full.log 2>&1
tail -3 gpurun_out/${TAG}_full.log
ls -la gpurun_out | grep ${TAG}
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'step frac',d['roofline_step']['frac'])
print('modes',d.get('modes'))
print('train',json.dumps(d.get('train'))[:1500])
"
