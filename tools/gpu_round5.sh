#!/bin/bash
TAG=${1:-r2g}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_conv_gemm.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${TAG}_ops.txt 2>&1
echo "ops+parity exit $?"; tail -3 gpurun_out/${TAG}_ops.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt
tail -6 gpurun_out/${TAG}_pytest.txt
timeout 900 python bench.py --steps 30 --warmup 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
VP3D_XPACK=0 timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-train --no-modes > gpurun_out/${TAG}_bench_noxpack.json 2>> gpurun_out/${TAG}_bench.err
VP3D_WAVE_NUM=3 timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-train --no-modes > gpurun_out/${TAG}_bench_wave3.json 2>> gpurun_out/${TAG}_bench.err
VP3D_WAVE_NUM=4 timeout 600 python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-cudnn --no-train --no-modes > gpurun_out/${TAG}_bench_wave4.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval.csv python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_prof.log 2>&1
VP3D_WAVE_NUM=3 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval_wave3.csv python tools/profile_steps.py eval fp16 >> gpurun_out/${TAG}_prof.log 2>&1
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
python -c "
import json
for f in ['bench','bench_noxpack','bench_wave3','bench_wave4']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),'dom frac',round(d['roofline']['frac'],3),'step frac',round(d['roofline_step']['frac'],3), 'launches', d['launches_per_step'])
        t=d.get('train') or {}; 
        if t: print('  train',t.get('ms_per_step'),t.get('ms_per_step_wall_incl_loss_item'),t.get('error'))
    except Exception as e: print(f,'ERR',e)
"
