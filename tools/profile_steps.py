#!/usr/bin/env python
"""Tiny driver for `ncu --metrics gpu__time_duration.sum`: a few eval forwards (given precision) or
training steps at the BASELINE shapes, bracketed by cudaProfilerStart/Stop so that only the last
iterations are captured (run ncu with --profile-from-start off).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_train.csv python tools/profile_steps.py train bf16
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import videopose3d_b200 as vp  # noqa: E402

ARC, C, J, F, N, T = [3, 3, 3, 3, 3], 1024, 17, 2, 1024, 243
what = sys.argv[1] if len(sys.argv) > 1 else "eval"
prec = sys.argv[2] if len(sys.argv) > 2 else ("mixed" if what == "eval" else "bf16")
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = (torch.rand(N, T, J, F, generator=g) * 2 - 1).to(dev)
tgt = (torch.randn(N, 1, J, 3, generator=g) * 0.3).to(dev)
if what == "eval":
    m = vp.TemporalModel(J, F, J, filter_widths=ARC, channels=C).to(dev).eval().set_precision(prec)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        m(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
else:
    from videopose3d_b200 import loss as vloss
    from videopose3d_b200.optim import FusedAdam
    m = vp.TemporalModelOptimized1f(J, F, J, filter_widths=ARC, dropout=0.25, channels=C).to(dev).train()
    m.set_train_precision(prec)
    opt = FusedAdam(m.parameters(), lr=1e-3, amsgrad=True)

    def step():   # the product path: fused loss head and optimiser (bench.py measure_train)
        opt.zero_grad()
        vloss.mpjpe(m(x), tgt).backward()
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done", what, prec)
