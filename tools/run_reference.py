#!/usr/bin/env python
"""Run the reference's `run.py` UNCHANGED on top of this package.

    cd <workdir with data/>            # e.g. written by tools/make_synthetic_h36m.py
    python tools/run_reference.py --reference /path/to/VideoPose3D [--swap model,loss,optim] \
        -- -k gt -arc 3,3,3,3,3 -e 1 ...        # everything after `--` goes to run.py

What it does: puts the reference checkout on sys.path, imports its `common.*` modules, replaces the
symbols `run.py` picks up with `from common.model import *` (run.py:21), `from common.loss import *`
(:22) and `optim.Adam` (:252, 264) by this package's drop-ins, then executes run.py as `__main__`.
No file of the reference is modified.
  model   TemporalModel, TemporalModelOptimized1f -> videopose3d_b200 (CUDA only)
  loss    mpjpe, weighted_mpjpe                   -> fused kernels (the NumPy metrics p_mpjpe /
          n_mpjpe / mean_velocity_error stay the reference's)
  optim   torch.optim.Adam                        -> FusedAdam
Default: model only.  The device-resident generators are not swapped in here: they yield CUDA
tensors where run.py expects NumPy arrays (`torch.from_numpy(batch.astype('float32'))`,
run.py:328-341, 402-406), so adopting them means deleting those lines -- see INTEGRATION.md.
"""
import argparse
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    if "--" in sys.argv:
        cut = sys.argv.index("--")
        own, passthrough = sys.argv[1:cut], sys.argv[cut + 1:]
    else:
        own, passthrough = sys.argv[1:], []
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--swap", default="model")
    args = ap.parse_args(own)
    swaps = set(filter(None, args.swap.split(",")))
    unknown = swaps - {"model", "loss", "optim"}
    if unknown:
        raise SystemExit(f"unknown --swap entries: {sorted(unknown)}")
    script = os.path.join(args.reference, "run.py")
    if not os.path.exists(script):
        raise SystemExit(f"{script} not found")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, args.reference)           # what `python run.py` would have as sys.path[0]

    import videopose3d_b200 as vp
    if "model" in swaps:
        import common.model as ref_model
        ref_model.TemporalModelBase = vp.TemporalModelBase
        ref_model.TemporalModel = vp.TemporalModel
        ref_model.TemporalModelOptimized1f = vp.TemporalModelOptimized1f
    if "loss" in swaps:
        import common.loss as ref_loss
        from videopose3d_b200 import loss as vloss
        ref_loss.mpjpe, ref_loss.weighted_mpjpe = vloss.mpjpe, vloss.weighted_mpjpe
    if "optim" in swaps:
        import torch.optim
        from videopose3d_b200.optim import FusedAdam
        torch.optim.Adam = FusedAdam
    sys.argv = [script] + passthrough
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
