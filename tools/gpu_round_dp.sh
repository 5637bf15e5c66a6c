#!/bin/bash
# multi-GPU session: bench.py under torchrun at N GPUs (eval replicas + data-parallel training block)
# with the reducer's overlap on / off and single-feature knobs for diagnosis
N=${1:-2}
TAG=${2:-r2e}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() {   # name, env assignments...
  local name=$1; shift
  env "$@" NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-cudnn --no-modes --no-cpu-baseline \
    > gpurun_out/${TAG}_bench_n${N}_${name}.json 2> gpurun_out/${TAG}_bench_n${N}_${name}.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_n${N}_${name}.json').read().strip().splitlines()[-1]); t=d.get('train_dp',{})
    print('${name}: eval',round(d['ms_per_step'],4),'train_dp',{k:(round(t[k],3) if isinstance(t.get(k),float) else t.get(k)) for k in ['ms_per_step','ms_per_step_wall_incl_loss_item','local_step_ms_no_collective','exposed_collective_ms','error']})
except Exception as e: print('${name} ERR',e)
"
}
run default VP3D_DUMMY=1
run default_2 VP3D_DUMMY=1
run default_3 VP3D_DUMMY=1
