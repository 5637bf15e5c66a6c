#!/bin/bash
# 2-GPU matrix of the data-parallel training block: overlap on/off, wire dtype, SM limit, PDL
N=${1:-2}
TAG=${2:-r2f}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --train-steps 20 --no-cudnn --no-modes --no-cpu-baseline $EXTRA \
    > gpurun_out/${TAG}_dp_${name}.json 2> gpurun_out/${TAG}_dp_${name}.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_dp_${name}.json').read().strip().splitlines()[-1]); t=d['train_dp']
    print('${name}', {k:(round(t[k],3) if isinstance(t.get(k),float) else t.get(k)) for k in ['ms_per_step','ms_per_step_wall_incl_loss_item','local_step_ms_no_collective','exposed_collective_ms','error']})
except Exception as e: print('${name}','ERR',e)
"
}
EXTRA=""
run overlap_fp32 VP3D_BENCH_DP_OVERLAP=1 NCCL_DEBUG=INFO
grep -E "NVLS|P2P|via|Channel 00|Connected|SHM|NET" gpurun_out/${TAG}_dp_overlap_fp32.err | head -12
run nooverlap_fp32 VP3D_BENCH_DP_OVERLAP=0
run overlap_sm132 VP3D_BENCH_DP_OVERLAP=1 VP3D_SM_LIMIT=132 NCCL_MAX_CTAS=16
run overlap_sm140_nopdl VP3D_BENCH_DP_OVERLAP=1 VP3D_SM_LIMIT=140 NCCL_MAX_CTAS=8 VP3D_PDL=0
run overlap_nopdl VP3D_BENCH_DP_OVERLAP=1 VP3D_PDL=0
EXTRA="--grad-wire bf16"
run nooverlap_bf16 VP3D_BENCH_DP_OVERLAP=0
run overlap_sm132_bf16 VP3D_BENCH_DP_OVERLAP=1 VP3D_SM_LIMIT=132 NCCL_MAX_CTAS=16
