#!/usr/bin/env python
"""Write a small synthetic Human3.6M-format dataset pair for smoke-testing `run.py` unchanged
(the real dataset is licensed and cannot ship).

    python tools/make_synthetic_h36m.py --reference /path/to/VideoPose3D --out /path/to/workdir/data
    cd /path/to/workdir && python /path/to/VideoPose3D/run.py -k gt -arc 3,3,3 -e 1 -b 128 ...

Files (formats documented at common/h36m_dataset.py:234-243 and data/prepare_data_h36m.py:151-171):
  data_3d_h36m.npz     positions_3d: {subject: {action: (frames, 32, 3) float32, world space, m}}
  data_2d_h36m_gt.npz  positions_2d: {subject: {action: [4 x (frames, 17, 2) float32, pixels]}},
                       metadata: {layout_name, num_joints, keypoints_symmetry}
The 3-D motion is a smooth random walk of a 32-joint cloud inside the capture volume; the 2-D
keypoints are its projection through the reference's own Human3.6M cameras (its camera code is
imported from --reference, exactly as data/prepare_data_h36m.py does), so the 2-D -> 3-D task is
consistent and a model can learn it.
"""
import argparse
import os
import sys

import numpy as np


def smooth_walk(rng, frames, dims, scale, smooth=25):
    steps = rng.normal(0, 1, (frames + smooth, dims))
    kernel = np.ones(smooth) / smooth
    path = np.stack([np.convolve(np.cumsum(steps[:, d]), kernel, mode="valid")[:frames]
                     for d in range(dims)], axis=1)
    path -= path.mean(axis=0)
    return path / (np.abs(path).max() + 1e-9) * scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of facebookresearch/VideoPose3D")
    ap.add_argument("--out", required=True, help="directory for the two .npz files")
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--actions", default="Walking,Directions 1")
    ap.add_argument("--subjects", default="S1,S5,S6,S7,S8,S9,S11")
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    from common.camera import image_coordinates, project_to_2d, world_to_camera
    from common.h36m_dataset import Human36mDataset
    from common.utils import wrap

    rng = np.random.RandomState(args.seed)
    os.makedirs(args.out, exist_ok=True)
    positions = {}
    for subject in args.subjects.split(","):
        positions[subject] = {}
        for action in args.actions.split(","):
            root = smooth_walk(rng, args.frames, 3, 0.8)
            root[:, 2] = 0.9 + 0.1 * root[:, 2]                      # hips about 0.9 m above the floor
            offsets = rng.uniform(-0.45, 0.45, (1, 32, 3)) * np.array([1.0, 1.0, 1.8])
            wobble = np.stack([smooth_walk(rng, args.frames, 3, 0.08) for _ in range(32)], axis=1)
            pose = root[:, None, :] + offsets + wobble
            pose[:, 0] = root                                         # joint 0 is the root
            positions[subject][action] = pose.astype(np.float32)
    path_3d = os.path.join(args.out, "data_3d_h36m.npz")
    np.savez_compressed(path_3d, positions_3d=positions)

    dataset = Human36mDataset(path_3d)                                # 32 -> 17 joints, cameras
    positions_2d = {}
    for subject in dataset.subjects():
        positions_2d[subject] = {}
        for action in dataset[subject].keys():
            anim = dataset[subject][action]
            views = []
            for cam in anim["cameras"]:
                pos_3d = world_to_camera(anim["positions"], R=cam["orientation"], t=cam["translation"])
                pos_2d = wrap(project_to_2d, pos_3d, cam["intrinsic"], unsqueeze=True)
                views.append(image_coordinates(pos_2d[..., :2], w=cam["res_w"], h=cam["res_h"])
                             .astype(np.float32))
            positions_2d[subject][action] = views
    metadata = {"layout_name": "h36m", "num_joints": dataset.skeleton().num_joints(),
                "keypoints_symmetry": [dataset.skeleton().joints_left(),
                                       dataset.skeleton().joints_right()]}
    np.savez_compressed(os.path.join(args.out, "data_2d_h36m_gt.npz"), positions_2d=positions_2d,
                        metadata=metadata)
    n = sum(len(a) for a in positions.values())
    print(f"wrote {path_3d} and data_2d_h36m_gt.npz: {len(positions)} subjects x "
          f"{n // len(positions)} actions x {args.frames} frames, 4 cameras")


if __name__ == "__main__":
    main()
