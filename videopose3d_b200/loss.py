"""Fused position losses (position part).

`mpjpe` and `weighted_mpjpe` with the reference's signatures and values (common/loss.py:11-17,
:19-25; used at run.py:359, 413, 452, 501): the loss and its gradient with respect to the
prediction come out of ONE launch (`vp3d_mpjpe_fwd_bwd`, csrc/step_ops.cu) instead of the
subtract / norm / mean kernels and their four backward kernels.  CUDA float32 only, no fallback.
"""
import torch

from . import _capi

__all__ = ["mpjpe", "weighted_mpjpe"]


class _Mpjpe(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predicted, target, weights):
        for t, what in ((predicted, "predicted"), (target, "target")):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise RuntimeError(f"videopose3d_b200.loss: {what} must be a CUDA float32 tensor "
                                   f"(got {t.device}, {t.dtype}); there is no fallback path")
        lib = _capi.load()
        pred = predicted.contiguous()
        tgt = target.contiguous()
        dims = pred.shape[-1]
        joints = pred.numel() // dims if dims else 0
        w = None
        if weights is not None:
            w = weights.to(device=pred.device, dtype=torch.float32).expand(pred.shape[:-1]).contiguous()
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        need_grad = ctx.needs_input_grad[0]
        dpred = torch.empty_like(pred) if need_grad else None
        with torch.cuda.device(pred.device):
            stream = torch.cuda.current_stream(pred.device).cuda_stream
            _capi.check(lib.vp3d_mpjpe_fwd_bwd(pred.data_ptr(), tgt.data_ptr(),
                                               w.data_ptr() if w is not None else None, joints, dims,
                                               loss.data_ptr(),
                                               dpred.data_ptr() if need_grad else None, stream),
                        "vp3d_mpjpe_fwd_bwd")
        ctx.dpred = dpred
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        dpred = ctx.dpred
        ctx.dpred = None
        return (dpred * grad_out if dpred is not None else None), None, None


def mpjpe(predicted, target):
    """Mean Euclidean distance between predicted and target joints (loss.py:11-17)."""
    assert predicted.shape == target.shape
    return _Mpjpe.apply(predicted, target, None)


def weighted_mpjpe(predicted, target, w):
    """Weighted mean Euclidean distance, `w` broadcast over the joint axis (loss.py:19-25)."""
    assert predicted.shape == target.shape
    assert w.shape[0] == predicted.shape[0]
    return _Mpjpe.apply(predicted, target, w)
