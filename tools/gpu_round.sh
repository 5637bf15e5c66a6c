#!/bin/bash
# One single-GPU box session: GPU test-suite, smoke, the bench line (both arms), launch lists of the
# eval forward and of the training step, in-kernel timeline, and one `ncu --set full` capture of the
# eval forward (-> profiles/traffic.json).  Everything lands in gpurun_out/<tag>_*.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_round.sh r2final'
TAG=${1:-r2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt
tail -4 gpurun_out/${TAG}_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 30 --warmup 8 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('bench value',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],3),'dom frac',round(d['roofline']['frac'],3),'step frac',round(d['roofline_step']['frac'],3),'launches',d['launches_per_step'])
print('  modes',{k:round(v['ms_per_step'],4) for k,v in d.get('modes',{}).items()})
t=d.get('train') or {}
print('  train',t.get('ms_per_step'),t.get('ms_per_step_wall_incl_loss_item'),t.get('error'),(t.get('cudnn_same_gpu') or {}))
print('  cudnn',{k:round(v['speedup_of_value'],1) for k,v in d['cudnn_same_gpu'].items() if isinstance(v,dict)})
r=json.loads(open('gpurun_out/${TAG}_bench_reference.json').read().strip().splitlines()[-1])
print('  reference arm',r.get('value'),r.get('cpu_baseline'))
"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval_fp16.csv python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_prof.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_train_bf16.csv python tools/profile_steps.py train bf16 >> gpurun_out/${TAG}_prof.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches_eval_fp16.csv 2>/dev/null | head -12
python tools/summarize_launches.py gpurun_out/${TAG}_launches_train_bf16.csv 2>/dev/null | grep -A12 totals
timeout 300 python tools/timeline.py fp16 > gpurun_out/${TAG}_timeline_eval_fp16.txt 2>&1
grep "^# rep" gpurun_out/${TAG}_timeline_eval_fp16.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/${TAG}_full_eval_fp16 -f python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_full.log 2>&1
tail -1 gpurun_out/${TAG}_full.log; ls -la gpurun_out/${TAG}_full_eval_fp16.ncu-rep
