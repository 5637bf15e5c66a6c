#!/bin/bash
# One GPU-box session: GPU test-suite, smoke, bench lines, launch list.  Everything lands in
# gpurun_out/<tag>_*.  Usage: tools/gpu_round.sh <tag> [quick]
TAG=${1:-r2}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.txt
tail -5 gpurun_out/${TAG}_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
tail -3 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
VP3D_PDL=0 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-cudnn > gpurun_out/${TAG}_bench_nopdl.json 2>> gpurun_out/${TAG}_bench.err
for prec in bf16 mixed bf16x3; do
  timeout 300 python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-cudnn --precision $prec > gpurun_out/${TAG}_bench_${prec}.json 2>> gpurun_out/${TAG}_bench.err
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches_eval.csv python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_prof.log 2>&1
python -c "
import json
for f in ['bench','bench_nopdl','bench_bf16','bench_mixed','bench_bf16x3']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
"
# one `ncu --set full` capture of every launch of one eval forward (default mode) for profiles/
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -o gpurun_out/${TAG}_full_eval_fp16 -f python tools/profile_steps.py eval fp16 > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out | tail -20
