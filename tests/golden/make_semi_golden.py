"""Golden vectors for the semi-supervised step (BASELINE configs[4]) and the training-size
gradient fixture (configs[2] shape), produced by the REAL reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_semi_golden.py

`semi_333_c64.npz` -- one step of run.py:345-396 executed literally with the reference's own
`TemporalModelOptimized1f` (position model J_out = 17 + trajectory model J_out = 1, as run.py:238-246
builds them), `common/loss.py` (mpjpe, weighted_mpjpe), `common/camera.py` (project_to_2d /
project_to_2d_linear) and the bone-length term with the Human3.6M skeleton parents
(`dataset.skeleton().parents()`, h36m_dataset.py:246-252).  Stored: all inputs, both state dicts, both model
outputs, every loss term, d(total)/d(model outputs) and every parameter gradient -- for the
distortion-aware and the linear projection.

`big_opt_33333_c1024_train.npz` -- TemporalModelOptimized1f arc 3,3,3,3,3, C = 1024 (the cfg3 shape),
N = 1024 windows (the full cfg3 batch), train mode, dropout 0: output, updated running statistics and the gradient of every
parameter; the 9 conv-weight gradients (up to 3.1 M elements each) are stored as a strided sample
of 4096 entries + their L2 norm + their sum, parameters are regenerated from the seed.

`causal_shift.json` -- total_causal_shift() / receptive_field() of both reference classes for a
list of architectures (model.py:41-61).
"""
import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from common.camera import project_to_2d, project_to_2d_linear  # noqa: E402  (the reference)
from common.h36m_dataset import h36m_cameras_intrinsic_params, h36m_skeleton  # noqa: E402
from common.camera import normalize_screen_coordinates  # noqa: E402
from common.loss import mpjpe, weighted_mpjpe  # noqa: E402
from common.model import TemporalModel, TemporalModelOptimized1f  # noqa: E402

from oracle import temporal_model_oracle as orc  # noqa: E402

COMMIT = "1afb1ca0f1237776518469876342fc8669d3f6a9"
SAMPLE = 4096


def h36m_parents_17():
    """parents() of the 17-joint skeleton exactly as Human36mDataset builds it."""
    sk = copy.deepcopy(h36m_skeleton)
    sk.remove_joints([4, 5, 9, 10, 11, 16, 20, 21, 22, 23, 24, 28, 29, 30, 31])
    sk._parents[11] = 8
    sk._parents[14] = 8
    return [int(p) for p in sk.parents()]


def h36m_intrinsics():
    """The four 9-vectors [f(2), c(2), k(3), p(2)] of h36m_dataset.py:218-231."""
    out = []
    for cam in copy.deepcopy(h36m_cameras_intrinsic_params):
        c = normalize_screen_coordinates(np.array(cam["center"], dtype="float32"), w=cam["res_w"],
                                         h=cam["res_h"]).astype("float32")
        f = np.array(cam["focal_length"], dtype="float32") / cam["res_w"] * 2
        out.append(np.concatenate((f, c, np.array(cam["radial_distortion"], dtype="float32"),
                                   np.array(cam["tangential_distortion"], dtype="float32"))))
    return np.stack(out).astype(np.float32)


def semi_case():
    arc, C, J, pad = [3, 3, 3], 64, 17, 13
    n_lab, n_unl, T = 12, 12, 27
    parents = h36m_parents_17()
    g = torch.Generator().manual_seed(2024)
    sd_pos = orc.make_state_dict(J, 2, J, arc, C, seed=11)
    sd_traj = orc.make_state_dict(J, 2, 1, arc, C, seed=12)
    # keep the predicted root trajectory in front of the camera for most samples (depth ~4 m) so
    # that both the clamped and the unclamped branch of camera.py:59 are exercised
    sd_traj["shrink.bias"] = torch.tensor([0.15, -0.1, 1.2])
    sd_pos["shrink.bias"] = sd_pos["shrink.bias"] * 0.5
    inputs_2d = torch.rand(n_lab, T, J, 2, generator=g) * 2 - 1
    inputs_2d_semi = torch.rand(n_unl, T, J, 2, generator=g) * 2 - 1
    inputs_3d = torch.randn(n_lab, 1, J, 3, generator=g) * 0.4
    inputs_3d[:, :, 0, 2] = torch.rand(n_lab, 1, generator=g) * 3 + 3         # root depth U(3, 6) m
    inputs_3d[:, :, 0, :2] = torch.randn(n_lab, 1, 2, generator=g) * 0.5
    cams = torch.from_numpy(h36m_intrinsics())
    cam_semi = cams[torch.randint(0, 4, (n_unl,), generator=g)].clone()
    flip = torch.rand(n_unl, generator=g) < 0.5                               # generators.py:146-152
    cam_semi[flip, 2] *= -1
    cam_semi[flip, 7] *= -1

    out = {}
    for linear in (False, True):
        model_pos = TemporalModelOptimized1f(J, 2, J, filter_widths=arc, dropout=0.0, channels=C)
        model_traj = TemporalModelOptimized1f(J, 2, 1, filter_widths=arc, dropout=0.0, channels=C)
        model_pos.load_state_dict(sd_pos)
        model_traj.load_state_dict(sd_traj)
        model_pos.train()
        model_traj.train()
        # ---- run.py:329-390, literally (skip = False, bone_length_term = True, no_proj = False)
        in3 = inputs_3d.clone()
        inputs_traj = in3[:, :, :1].clone()
        in3[:, :, 0] = 0
        split_idx = in3.shape[0]
        inputs_2d_cat = torch.cat((inputs_2d, inputs_2d_semi), dim=0)
        predicted_3d_pos_cat = model_pos(inputs_2d_cat)
        predicted_3d_pos_cat.retain_grad()
        loss_3d_pos = mpjpe(predicted_3d_pos_cat[:split_idx], in3)
        loss_total = loss_3d_pos
        predicted_traj_cat = model_traj(inputs_2d_cat)
        predicted_traj_cat.retain_grad()
        w = 1 / inputs_traj[:, :, :, 2]
        loss_traj = weighted_mpjpe(predicted_traj_cat[:split_idx], inputs_traj, w)
        loss_total = loss_total + loss_traj
        predicted_semi = predicted_3d_pos_cat[split_idx:]
        target_semi = inputs_2d_semi[:, pad:-pad, :, :2].contiguous()
        projection_func = project_to_2d_linear if linear else project_to_2d
        reconstruction_semi = projection_func(predicted_semi + predicted_traj_cat[split_idx:], cam_semi)
        loss_reconstruction = mpjpe(reconstruction_semi, target_semi)
        loss_total = loss_total + loss_reconstruction
        dists = predicted_3d_pos_cat[:, :, 1:] - predicted_3d_pos_cat[:, :, parents[1:]]
        bone_lengths = torch.mean(torch.norm(dists, dim=3), dim=1)
        penalty = torch.mean(torch.abs(torch.mean(bone_lengths[:split_idx], dim=0)
                                       - torch.mean(bone_lengths[split_idx:], dim=0)))
        loss_total = loss_total + penalty
        loss_total.backward()
        tag = "lin/" if linear else "full/"
        out[tag + "pred_pos"] = predicted_3d_pos_cat.detach().numpy()
        out[tag + "pred_traj"] = predicted_traj_cat.detach().numpy()
        out[tag + "losses"] = np.array([loss_3d_pos.item(), loss_traj.item(), loss_reconstruction.item(),
                                        penalty.item(), loss_total.item()], dtype=np.float64)
        out[tag + "reconstruction"] = reconstruction_semi.detach().numpy()
        out[tag + "d_pred_pos"] = predicted_3d_pos_cat.grad.numpy()
        out[tag + "d_pred_traj"] = predicted_traj_cat.grad.numpy()
        for name, m in (("pos", model_pos), ("traj", model_traj)):
            for k, prm in m.named_parameters():
                out[f"{tag}grad_{name}/{k}"] = prm.grad.numpy()
        frac_clamped = float(((predicted_semi + predicted_traj_cat[split_idx:])[..., :2].abs()
                              >= (predicted_semi + predicted_traj_cat[split_idx:])[..., 2:].abs()).float().mean())
    meta = dict(arc=arc, C=C, J=J, pad=pad, n_labeled=n_lab, n_unlabeled=n_unl, T=T, parents=parents,
                frac_clamped=frac_clamped, torch=torch.__version__, reference_commit=COMMIT,
                sequence="run.py:329-390 (skip=False, bone_length_term=True)")
    out["meta"] = np.array(json.dumps(meta))
    out["inputs_2d"] = inputs_2d.numpy()
    out["inputs_2d_semi"] = inputs_2d_semi.numpy()
    out["inputs_3d"] = inputs_3d.numpy()
    out["cam_semi"] = cam_semi.numpy()
    for name, sd in (("pos", sd_pos), ("traj", sd_traj)):
        for k, v in sd.items():
            out[f"sd_{name}/{k}"] = v.numpy()
    return out


def sample_idx(numel):
    if numel <= SAMPLE:
        return np.arange(numel)
    return np.linspace(0, numel - 1, SAMPLE).astype(np.int64)


def cfg3_train_case():
    arc, C, J, N, T, seed, momentum = [3, 3, 3, 3, 3], 1024, 17, 1024, 243, 4321, 0.1
    sd = orc.make_state_dict(J, 2, J, arc, C, seed=seed)
    x = orc.make_input(N, T, J, 2, seed=seed + 1)
    model = TemporalModelOptimized1f(J, 2, J, filter_widths=arc, dropout=0.0, channels=C)
    model.load_state_dict(sd)
    model.train()
    model.set_bn_momentum(momentum)
    y = model(x)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed + 2))
    (y * gy).sum().backward()
    out = {"y": y.detach().numpy(), "gy": gy.numpy()}
    for k, prm in model.named_parameters():
        g = prm.grad.reshape(-1).numpy()
        idx = sample_idx(g.size)
        out["gidx/" + k] = idx
        out["gval/" + k] = g[idx]
        out["gnorm/" + k] = np.array([np.sqrt((g.astype(np.float64) ** 2).sum()), g.astype(np.float64).sum(),
                                      np.abs(g).max()])
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["new/" + k] = v.numpy()
    # parameter checksums: the test regenerates the parameters from the seed and must hit these
    out["sd_check"] = np.array([float(v.double().sum()) for k, v in sorted(sd.items())])
    meta = dict(cls="TemporalModelOptimized1f", J=J, F=2, Jout=J, fw=arc, C=C, N=N, T=T, seed=seed,
                momentum=momentum, causal=False, dense=False, torch=torch.__version__,
                reference_commit=COMMIT, sample=SAMPLE)
    out["meta"] = np.array(json.dumps(meta))
    return out


def causal_shift_table():
    rows = []
    for arc in ([3], [3, 3], [3, 3, 3], [3, 3, 3, 3, 3], [3, 5, 3], [5, 3, 7], [1, 3], [7, 7]):
        for causal in (False, True):
            for cls in (TemporalModel, TemporalModelOptimized1f):
                m = cls(17, 2, 17, filter_widths=arc, causal=causal, channels=8)
                rows.append(dict(cls=cls.__name__, arc=arc, causal=causal,
                                 total_causal_shift=int(m.total_causal_shift()),
                                 receptive_field=int(m.receptive_field()),
                                 pad=[int(p) for p in m.pad],
                                 causal_shift=[int(s) for s in m.causal_shift]))
    return rows


def main():
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])
    if not only or "semi" in only:
        data = semi_case()
        path = os.path.join(HERE, "semi_333_c64.npz")
        np.savez_compressed(path, **data)
        print("semi_333_c64:", json.loads(str(data["meta"]))["frac_clamped"], "clamped;",
              {k: data[k].tolist() for k in ("full/losses", "lin/losses")},
              f"-> {os.path.getsize(path) / 1024:.0f} KiB")
    if not only or "cfg3" in only:
        data = cfg3_train_case()
        path = os.path.join(HERE, "big_opt_33333_c1024_train.npz")
        np.savez_compressed(path, **data)
        print(f"opt_33333_c1024_train: y{tuple(data['y'].shape)} -> {os.path.getsize(path) / 1024:.0f} KiB")
    if not only or "shift" in only:
        with open(os.path.join(HERE, "causal_shift.json"), "w") as f:
            json.dump(causal_shift_table(), f, indent=0)
        print("causal_shift.json written")


if __name__ == "__main__":
    main()
