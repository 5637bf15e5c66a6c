#!/bin/bash
TAG=${1:-r2d}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -m gpu -q -x > gpurun_out/${TAG}_ops.txt 2>&1
echo "ops exit $?"; tail -3 gpurun_out/${TAG}_ops.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt
tail -8 gpurun_out/${TAG}_pytest.txt
timeout 900 python bench.py --steps 30 --warmup 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/${TAG}_full_eval_fp16 -f python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_full.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_train.csv python tools/profile_steps.py train bf16 > gpurun_out/${TAG}_prof.log 2>&1
python -c "
import json
for f in ['bench']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'dom frac',d['roofline']['frac'],'step frac',d['roofline_step']['frac'])
        print('  modes',{k:round(v['ms_per_step'],4) for k,v in d.get('modes',{}).items()})
        t=d.get('train',{}); print('  train',t.get('ms_per_step'),t.get('error'), (t.get('cudnn_same_gpu') or {}).get('fp32_tf32'))
    except Exception as e: print(f,'ERR',e)
"
