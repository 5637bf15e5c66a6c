#!/usr/bin/env python
"""Secondary measurements that bench.py's single JSON line does not carry (1 GPU):

  * the number north_star asks to beat: the reference network executed by stock PyTorch/cuDNN on the
    same B200 (fp32 with TF32 convolutions = PyTorch default, TF32 off, bf16 autocast), for the
    eval forward of BASELINE configs[1] and the training step of configs[2];
  * this repo's eval forward in all three precision modes and its Optimized1f training step
    (forward + backward + Adam(amsgrad) as in run.py:252, 409-420).

The cuDNN baseline is built here from plain torch.nn modules following common/model.py:85-138 /
151-197 (it is a measurement target, not the product and not the oracle).  CUDA-event timing,
10 warm-up + N timed iterations, cudnn.benchmark on, GPU-resident synthetic inputs.

    python tools/bench_extra.py [--iters 30] [--what eval,train] > gpurun_out/extra.jsonl
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import videopose3d_b200 as vp  # noqa: E402

ARC, C, J, F, N, T = [3, 3, 3, 3, 3], 1024, 17, 2, 1024, 243


class CudnnTemporal(nn.Module):
    """Stock-PyTorch execution of the reference architecture (dilated or strided)."""

    def __init__(self, strided, dropout=0.25):
        super().__init__()
        self.strided = strided
        fw = ARC
        self.expand = nn.Conv1d(J * F, C, fw[0], stride=fw[0] if strided else 1, bias=False)
        self.expand_bn = nn.BatchNorm1d(C, momentum=0.1)
        convs, bns = [], []
        d = fw[0]
        self.pads = []
        for w in fw[1:]:
            self.pads.append((w - 1) * d // 2)
            convs.append(nn.Conv1d(C, C, w, stride=w, bias=False) if strided
                         else nn.Conv1d(C, C, w, dilation=d, bias=False))
            bns.append(nn.BatchNorm1d(C, momentum=0.1))
            convs.append(nn.Conv1d(C, C, 1, bias=False))
            bns.append(nn.BatchNorm1d(C, momentum=0.1))
            d *= w
        self.convs, self.bns = nn.ModuleList(convs), nn.ModuleList(bns)
        self.shrink = nn.Conv1d(C, J * 3, 1)
        self.drop, self.relu = nn.Dropout(dropout), nn.ReLU(inplace=True)

    def forward(self, x):
        n = x.shape[0]
        x = x.view(n, x.shape[1], -1).permute(0, 2, 1)
        x = self.drop(self.relu(self.expand_bn(self.expand(x))))
        for i, w in enumerate(ARC[1:]):
            if self.strided:
                res = x[:, :, w // 2:: w]
            else:
                p = self.pads[i]
                res = x[:, :, p: x.shape[2] - p]
            x = self.drop(self.relu(self.bns[2 * i](self.convs[2 * i](x))))
            x = res + self.drop(self.relu(self.bns[2 * i + 1](self.convs[2 * i + 1](x))))
        x = self.shrink(x)
        return x.permute(0, 2, 1).reshape(n, -1, J, 3)


def timeit(fn, iters, warmup=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2], ms[0]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--what", default="eval,train")
    args = ap.parse_args()
    what = set(args.what.split(","))
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(N, T, J, F, generator=g) * 2 - 1).to(dev)
    tgt = (torch.randn(N, 1, J, 3, generator=g) * 0.3).to(dev)

    if "eval" in what:
        m = vp.TemporalModel(J, F, J, filter_widths=ARC, channels=C).to(dev).eval()
        for prec in ("mixed", "bf16", "bf16x3"):
            m.set_precision(prec)
            with torch.no_grad():
                med, best = timeit(lambda: m(x), args.iters)
            emit(what="eval_forward", impl="vp3d_b200", precision=prec, ms_median=med, ms_best=best,
                 frames_per_s=N / med * 1e3, launches=m.last_launch_count())
        del m
        ref = CudnnTemporal(strided=False).to(dev).eval()
        for name, tf32, autocast in (("fp32_tf32", True, False), ("fp32_ieee", False, False),
                                     ("bf16_autocast", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32

            def run():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    return ref(x)
            med, best = timeit(run, max(5, args.iters // 3), warmup=3)
            emit(what="eval_forward", impl="pytorch_cudnn_reference_arch", precision=name,
                 ms_median=med, ms_best=best, frames_per_s=N / med * 1e3)
        del ref
        torch.cuda.empty_cache()

    if "seq" in what:
        # the reference's inference use (UnchunkedGenerator + evaluate(), run.py:652-721): one whole
        # video per forward, test-time flip augmentation -> batch of 2 sequences, every frame
        # predicted.  This is the dilated schedule (no cone pruning possible).
        frames = 6000
        xs = (torch.rand(2, frames + 242, J, F, generator=g) * 2 - 1).to(dev)
        m = vp.TemporalModel(J, F, J, filter_widths=ARC, channels=C).to(dev).eval()
        for prec in ("mixed", "bf16", "bf16x3"):
            m.set_precision(prec)
            with torch.no_grad():
                med, best = timeit(lambda: m(xs), args.iters)
            emit(what="eval_sequence_2x6000", impl="vp3d_b200", precision=prec, ms_median=med,
                 ms_best=best, frames_per_s=2 * frames / med * 1e3, launches=m.last_launch_count())
        del m
        ref = CudnnTemporal(strided=False).to(dev).eval()
        for name, tf32, autocast in (("fp32_tf32", True, False), ("bf16_autocast", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32

            def run_seq():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    return ref(xs)
            med, best = timeit(run_seq, max(5, args.iters // 3), warmup=3)
            emit(what="eval_sequence_2x6000", impl="pytorch_cudnn_reference_arch", precision=name,
                 ms_median=med, ms_best=best, frames_per_s=2 * frames / med * 1e3)
        del ref
        torch.cuda.empty_cache()

    if "train" in what:
        for prec in ("bf16", "bf16x3"):
            m = vp.TemporalModelOptimized1f(J, F, J, filter_widths=ARC, channels=C).to(dev).train()
            m.set_train_precision(prec)
            opt = torch.optim.Adam(m.parameters(), lr=1e-3, amsgrad=True)

            def step():
                opt.zero_grad()
                loss = torch.mean(torch.norm(m(x) - tgt, dim=-1))
                loss.backward()
                opt.step()
            med, best = timeit(step, args.iters)

            def fwd_bwd():
                opt.zero_grad()
                torch.mean(torch.norm(m(x) - tgt, dim=-1)).backward()
            med_fb, _ = timeit(fwd_bwd, args.iters, warmup=3)
            emit(what="train_step", impl="vp3d_b200", precision=prec, ms_median=med, ms_best=best,
                 ms_fwd_bwd=med_fb, frames_per_s=N / med * 1e3)
            del m, opt
            torch.cuda.empty_cache()
        for name, tf32, autocast in (("fp32_tf32", True, False), ("bf16_autocast", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ref = CudnnTemporal(strided=True).to(dev).train()
            opt = torch.optim.Adam(ref.parameters(), lr=1e-3, amsgrad=True)

            def step():
                opt.zero_grad()
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    out = ref(x)
                loss = torch.mean(torch.norm(out.float() - tgt, dim=-1))
                loss.backward()
                opt.step()
            med, best = timeit(step, args.iters)
            emit(what="train_step", impl="pytorch_cudnn_reference_arch", precision=name, ms_median=med,
                 ms_best=best, frames_per_s=N / med * 1e3)
            del ref, opt
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
