#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt
tail -6 gpurun_out/${TAG}_pytest.txt
grep -h "cfg3-shape\|run-to-run\|dropout keep\|semi-supervised step\|epoch 2\|final:" gpurun_out/${TAG}_pytest.txt | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 30 --warmup 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval.csv python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_prof.log 2>&1
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
python -c "
import json
for f in ['bench']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),'dom frac',round(d['roofline']['frac'],3),'step frac',round(d['roofline_step']['frac'],3), 'launches', d['launches_per_step'])
        print('  modes',{k:round(v['ms_per_step'],4) for k,v in d.get('modes',{}).items()})
        t=d.get('train') or {}; 
        if t: print('  train',t.get('ms_per_step'),t.get('ms_per_step_wall_incl_loss_item'),t.get('error'))
        print('  cudnn', {k:round(v['speedup_of_value'],1) for k,v in d['cudnn_same_gpu'].items() if isinstance(v,dict)})
    except Exception as e: print(f,'ERR',e)
"
