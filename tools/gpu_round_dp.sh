#!/bin/bash
# multi-GPU session: bench.py under torchrun at N GPUs (eval replicas + data-parallel training block)
N=${1:-2}
TAG=${2:-r2e}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
for wire in fp32 bf16; do
  EXTRA=""; [ "$wire" = "bf16" ] && EXTRA="--grad-wire bf16"
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-cudnn --no-modes $EXTRA \
    > gpurun_out/${TAG}_bench_n${N}_${wire}.json 2> gpurun_out/${TAG}_bench_n${N}_${wire}.err
  echo "n=$N wire=$wire exit $?"; tail -c 400 gpurun_out/${TAG}_bench_n${N}_${wire}.err
done
python -c "
import json
for w in ['fp32','bf16']:
    try:
        d=json.loads(open('gpurun_out/${TAG}_bench_n${N}_%s.json'%w).read().strip().splitlines()[-1])
        t=d.get('train_dp',{})
        print(w,'eval value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
        print('   train_dp',{k:t.get(k) for k in ['ms_per_step','local_step_ms_no_collective','exposed_collective_ms','weak_scaling_efficiency_vs_local_step','frames_per_s','allreduce_bytes_per_step','allreduce_slices_per_step','error']})
    except Exception as e: print(w,'ERR',e)
"
