#!/bin/bash
# e2e (host-buffer API) with different H2D chunk counts; prints ms_per_step of the e2e leg
for c in 1 2 4; do
  VP3D_HOST_CHUNKS=$c python bench.py --no-cpu-baseline --steps 100 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('chunks', $c, 'e2e ms/step', round(d['e2e']['ms_per_step'],4), 'frames/s', round(d['e2e']['value']), 'device-resident ms', round(d['ms_per_step'],4))"
done
