#!/bin/bash
# GPU session: conv GEMM op tests with and without CTA pairs (separate processes: a trap poisons
# the context), then the whole suite, benches (pair on/off), launch lists.
TAG=${1:-r2c}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
VP3D_PAIR=0 timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -m gpu -q -x > gpurun_out/${TAG}_ops_nopair.txt 2>&1
echo "ops nopair exit $?"; tail -3 gpurun_out/${TAG}_ops_nopair.txt
VP3D_PAIR=1 timeout 600 python -m pytest tests/test_gpu_conv_gemm.py -m gpu -q > gpurun_out/${TAG}_ops_pair.txt 2>&1
PAIR_RC=$?
echo "ops pair exit $PAIR_RC"; tail -5 gpurun_out/${TAG}_ops_pair.txt
if [ $PAIR_RC -ne 0 ]; then export VP3D_PAIR=0; echo "PAIR DISABLED for the rest of this session"; fi
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt
tail -8 gpurun_out/${TAG}_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 30 --warmup 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
VP3D_PAIR=0 timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn > gpurun_out/${TAG}_bench_nopair.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval.csv python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_prof.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_train.csv python tools/profile_steps.py train bf16 >> gpurun_out/${TAG}_prof.log 2>&1
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
python -c "
import json
for f in ['bench','bench_nopair']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',d['value'],d['ms_per_step'],'e2e',d['e2e']['value'],'dom frac',d['roofline']['frac'],'step frac',d['roofline_step']['frac'])
        print('  modes',{k:round(v['ms_per_step'],4) for k,v in d.get('modes',{}).items()})
        t=d.get('train',{}); print('  train',t.get('ms_per_step'),t.get('error'), (t.get('cudnn_same_gpu') or {}).get('fp32_tf32'))
    except Exception as e: print(f,'ERR',e)
"
