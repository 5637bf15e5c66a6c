"""Fused losses.

`mpjpe` and `weighted_mpjpe` with the reference's signatures and values (common/loss.py:11-17,
:19-25; used at run.py:359, 413, 452, 501): the loss and its gradient with respect to the
prediction come out of ONE launch (`vp3d_mpjpe_fwd_bwd`, csrc/step_ops.cu) instead of the
subtract / norm / mean kernels and their four backward kernels.  `projected_mpjpe` is the
re-projection loss of the semi-supervised branch (run.py:374-379): camera projection of
`predicted_pos + predicted_traj` (common/camera.py:37-88) and the 2-D mpjpe, with the gradients for
both model outputs, in one launch (`vp3d_projected_mpjpe_fwd_bwd`).  CUDA float32 only, no
fallback.  `bone_length_penalty` (run.py:385-390) is composed from torch ops.
"""
import torch

from . import _capi

__all__ = ["mpjpe", "weighted_mpjpe", "projected_mpjpe", "bone_length_penalty"]


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
        # kept for the lifetime of the graph: a second backward (retain_graph=True, or the loss
        # feeding two backward passes) gets the same gradient again, as with the torch expression
        ctx.dpred = dpred
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        dpred = ctx.dpred
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


class _ProjectedMpjpe(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pos, traj, cam, target, linear):
        for t, what in ((pos, "predicted_pos"), (traj, "predicted_traj"), (cam, "camera_params"),
                        (target, "target_2d")):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise RuntimeError(f"videopose3d_b200.loss: {what} must be a CUDA float32 tensor "
                                   f"(got {t.device}, {t.dtype}); there is no fallback path")
        n, frames, joints = pos.shape[0], pos.shape[1], pos.shape[2]
        assert pos.dim() == 4 and pos.shape[-1] == 3
        assert traj.shape == (n, frames, 1, 3), traj.shape
        assert cam.shape == (n, 9), cam.shape                      # camera.py:47-49
        assert target.shape == (n, frames, joints, 2), target.shape
        lib = _capi.load()
        pos_c, traj_c, cam_c, tgt_c = (t.contiguous() for t in (pos, traj, cam, target))
        loss = torch.empty((), dtype=torch.float32, device=pos.device)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dpos = torch.empty_like(pos_c) if need_grad else None
        dtraj = torch.empty_like(traj_c) if need_grad else None
        with torch.cuda.device(pos.device):
            stream = torch.cuda.current_stream(pos.device).cuda_stream
            _capi.check(lib.vp3d_projected_mpjpe_fwd_bwd(
                pos_c.data_ptr(), traj_c.data_ptr(), cam_c.data_ptr(), tgt_c.data_ptr(), n, frames,
                joints, int(bool(linear)), loss.data_ptr(),
                dpos.data_ptr() if need_grad else None, dtraj.data_ptr() if need_grad else None,
                stream), "vp3d_projected_mpjpe_fwd_bwd")
        ctx.grads = (dpos, dtraj)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        dpos, dtraj = ctx.grads
        if dpos is None:
            return None, None, None, None, None
        return dpos * grad_out, dtraj * grad_out, None, None, None


def projected_mpjpe(predicted_pos, predicted_traj, camera_params, target_2d, linear=False):
    """`mpjpe(project_to_2d(predicted_pos + predicted_traj, camera_params), target_2d)` -- the
    reconstruction loss of run.py:374-379 (`project_to_2d_linear` when `linear`)."""
    return _ProjectedMpjpe.apply(predicted_pos, predicted_traj, camera_params, target_2d, linear)


def bone_length_penalty(predicted_3d_pos_cat, split_idx, parents):
    """Kinematic term of the semi-supervised branch (run.py:385-390): mean absolute difference
    between the per-bone mean lengths of the labelled rows `[:split_idx]` and the unlabelled rows
    `[split_idx:]`.  `parents`: the skeleton's parent index per joint (joint 0 is the root).
    A handful of tiny reductions on (N, T, J, 3): composed from torch ops, no custom kernel."""
    parents = list(parents)
    dists = predicted_3d_pos_cat[:, :, 1:] - predicted_3d_pos_cat[:, :, parents[1:]]
    bone_lengths = torch.mean(torch.norm(dists, dim=3), dim=1)
    return torch.mean(torch.abs(torch.mean(bone_lengths[:split_idx], dim=0)
                                - torch.mean(bone_lengths[split_idx:], dim=0)))
