#!/bin/bash
# scaling session on an N-GPU box: bench.py exactly as the driver launches it
NMAX=${1:-8}
TAG=${2:-r2i}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for N in $NMAX; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 \
    > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
  echo "N=$N exit $?"; tail -c 300 gpurun_out/${TAG}_bench_n${N}.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_n${N}.json').read().strip().splitlines()[-1]); t=d.get('train_dp',{})
    print('N=$N eval',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']))
    print('  train_dp',{k:(round(t[k],3) if isinstance(t.get(k),float) else t.get(k)) for k in ['ms_per_step','ms_per_step_wall_incl_loss_item','local_step_ms_no_collective','exposed_collective_ms','weak_scaling_efficiency_vs_local_step','frames_per_s','error']})
except Exception as e: print('N=$N ERR',e)
"
done
