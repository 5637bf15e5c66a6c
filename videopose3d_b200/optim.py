"""Single-launch Adam / AMSGrad.

Drop-in for the optimiser the reference constructs with
`optim.Adam(model.parameters(), lr=lr, amsgrad=True)` (run.py:252, 264) and drives with
`optimizer.zero_grad()` / `optimizer.step()` (run.py:347, 396, 409, 420) and
`param_group['lr'] *= lr_decay` (run.py:583-586): same constructor arguments, same update rule,
same `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq`, `max_exp_avg_sq` per parameter), so
checkpoints written by either optimiser load into the other (run.py:295-296, 600-608).

`step()` hands every parameter of a group to `vp3d_adam_step` (csrc/step_ops.cu): one kernel launch
reads p, g, m, v, vmax and writes p, m, v, vmax once (36 B per element) instead of the eight
multi-tensor launches of torch's implementation.  CUDA float32 contiguous tensors only; anything
else raises -- there is no fallback path.
"""
import ctypes

import torch

from . import _capi

__all__ = ["FusedAdam"]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("amsgrad", False)
        for st in self.state.values():  # checkpoints from old torch versions keep `step` as an int
            if "step" in st and not torch.is_tensor(st["step"]):
                st["step"] = torch.tensor(float(st["step"]), dtype=torch.float32)

    @staticmethod
    def _check(t, what):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f"FusedAdam: {what} must be a contiguous CUDA float32 tensor "
                               f"(got {t.device}, {t.dtype}, contiguous={t.is_contiguous()}); "
                               "there is no fallback path")

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _capi.load()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            by_step = {}  # parameters that share a step count share a launch
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                self._check(p, "parameter")
                self._check(p.grad, "gradient")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["amsgrad"]:
                        st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                elif group["amsgrad"] and "max_exp_avg_sq" not in st:
                    st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                for key in ("exp_avg", "exp_avg_sq") + (("max_exp_avg_sq",) if group["amsgrad"] else ()):
                    self._check(st[key], key)
                st["step"] += 1
                by_step.setdefault(int(st["step"].item()), []).append((p, st))
            for step, items in by_step.items():
                dev = items[0][0].device
                if any(p.device != dev for p, _ in items):
                    raise RuntimeError("FusedAdam: parameters of one group must share a device")
                table = (_capi.AdamTensor * len(items))()
                for row, (p, st) in zip(table, items):
                    row.param = p.data_ptr()
                    row.grad = p.grad.data_ptr()
                    row.exp_avg = st["exp_avg"].data_ptr()
                    row.exp_avg_sq = st["exp_avg_sq"].data_ptr()
                    row.max_exp_avg_sq = st["max_exp_avg_sq"].data_ptr() if group["amsgrad"] else None
                    row.numel = p.numel()
                with torch.cuda.device(dev):
                    stream = torch.cuda.current_stream(dev).cuda_stream
                    _capi.check(lib.vp3d_adam_step(table, len(items), step, float(group["lr"]),
                                                   float(beta1), float(beta2), float(group["eps"]),
                                                   float(group["weight_decay"]), stream),
                                "vp3d_adam_step")
                for p, _ in items:
                    # the kernel wrote through raw pointers: tell autograd / the weight-pack cache
                    torch.autograd.graph.increment_version(p)
        return loss
