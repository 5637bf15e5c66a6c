"""Golden batches from the REAL reference generators.

Run in the build container only (needs /root/reference):

    python tests/golden/make_generator_golden.py

Imports `common.generators` from /root/reference unchanged, feeds it small seeded synthetic
datasets and stores the datasets, the constructor arguments and every batch the reference yields as
`gen_*.npz` fixtures.  tests/test_generator_oracle.py pins oracle/generator_oracle.py to them and
tests/test_gpu_generators.py checks the device generators against both.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

from common.generators import ChunkedGenerator, UnchunkedGenerator  # noqa: E402  (the reference)

# Human3.6M 17-joint symmetry (reference data/data_utils.py:28-35, h36m skeleton joints_left/right)
KPS_L, KPS_R = [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]

CASES = {
    # supervised training stream: shuffled + mirrored, chunk_length 1, pad 13 (arc 3,3,3), 2 epochs
    "gen_sup_shuffle_aug": dict(kind="chunked", lengths=[17, 40, 9, 23], J=17, F=2, batch_size=32,
                                use_cam=False, use_3d=True, chunk_length=1, pad=13, causal_shift=0,
                                shuffle=True, random_seed=1234, augment=True, epochs=2),
    # semi-supervised style: cameras, no 3-D, causal padding, fixed order
    "gen_cam_causal_noshuffle": dict(kind="chunked", lengths=[12, 30, 5], J=17, F=2, batch_size=20,
                                     use_cam=True, use_3d=False, chunk_length=1, pad=4,
                                     causal_shift=4, shuffle=False, random_seed=7, augment=True,
                                     epochs=1),
    # longer chunks: the chunk grid is centred, first chunk starts before frame 0
    "gen_chunk4": dict(kind="chunked", lengths=[10, 21, 7], J=16, F=3, batch_size=5, use_cam=True,
                       use_3d=True, chunk_length=4, pad=3, causal_shift=0, shuffle=True,
                       random_seed=99, augment=False, epochs=2),
    # 2-D only, no augmentation, short last batch
    "gen_2d_only": dict(kind="chunked", lengths=[15, 8], J=17, F=2, batch_size=10, use_cam=False,
                        use_3d=False, chunk_length=1, pad=1, causal_shift=0, shuffle=True,
                        random_seed=5, augment=False, epochs=1),
    # endless stream: 2.5 epochs worth of batches through one iterator
    "gen_endless": dict(kind="chunked", lengths=[11, 14], J=17, F=2, batch_size=8, use_cam=False,
                        use_3d=True, chunk_length=1, pad=2, causal_shift=0, shuffle=True,
                        random_seed=11, augment=True, endless=True, take_batches=17),
    # evaluation generators
    "gen_unchunked_aug": dict(kind="unchunked", lengths=[19, 6, 33], J=17, F=2, use_cam=True,
                              use_3d=True, pad=13, causal_shift=0, augment=True),
    "gen_unchunked_causal": dict(kind="unchunked", lengths=[8, 25], J=17, F=2, use_cam=False,
                                 use_3d=True, pad=4, causal_shift=4, augment=False),
}


def make_dataset(cfg, seed):
    rng = np.random.RandomState(seed)
    p2 = [rng.uniform(-1, 1, (n, cfg["J"], cfg["F"])).astype(np.float32) for n in cfg["lengths"]]
    p3 = ([rng.normal(0, 0.5, (n, cfg["J"], 3)).astype(np.float32) for n in cfg["lengths"]]
          if cfg["use_3d"] else None)
    cams = ([rng.uniform(-1, 1, 9).astype(np.float32) for _ in cfg["lengths"]]
            if cfg["use_cam"] else None)
    return cams, p3, p2


def sym(cfg):
    if cfg["J"] == 17:
        return KPS_L, KPS_R
    return [1, 2, 3, 10, 11], [4, 5, 6, 13, 14]  # an arbitrary symmetric split for other skeletons


def build(name, cfg):
    seed = sum(ord(c) for c in name)
    cams, p3, p2 = make_dataset(cfg, seed)
    left, right = sym(cfg)
    batches = []
    if cfg["kind"] == "chunked":
        gen = ChunkedGenerator(cfg["batch_size"], cams, p3, p2, cfg["chunk_length"], pad=cfg["pad"],
                               causal_shift=cfg["causal_shift"], shuffle=cfg["shuffle"],
                               random_seed=cfg["random_seed"], augment=cfg["augment"], kps_left=left,
                               kps_right=right, joints_left=left, joints_right=right,
                               endless=cfg.get("endless", False))
        if cfg.get("endless"):
            it = gen.next_epoch()
            for _ in range(cfg["take_batches"]):
                c, b3, b2 = next(it)
                batches.append((None if c is None else c.copy(), None if b3 is None else b3.copy(),
                                b2.copy()))
        else:
            for _ in range(cfg["epochs"]):
                for c, b3, b2 in gen.next_epoch():
                    batches.append((None if c is None else c.copy(),
                                    None if b3 is None else b3.copy(), b2.copy()))
        extra = dict(num_frames=gen.num_frames())
    else:
        gen = UnchunkedGenerator(cams, p3, p2, pad=cfg["pad"], causal_shift=cfg["causal_shift"],
                                 augment=cfg["augment"], kps_left=left, kps_right=right,
                                 joints_left=left, joints_right=right)
        for c, b3, b2 in gen.next_epoch():
            batches.append((c, b3, b2))
        extra = dict(num_frames=gen.num_frames())
    out = {"config": np.array(json.dumps(dict(cfg, left=left, right=right, seed=seed, **extra))),
           "n_batches": np.array(len(batches))}
    for i, a in enumerate(p2):
        out[f"p2_{i}"] = a
    if p3 is not None:
        for i, a in enumerate(p3):
            out[f"p3_{i}"] = a
    if cams is not None:
        out["cams"] = np.stack(cams)
    for i, (c, b3, b2) in enumerate(batches):
        out[f"b2_{i}"] = b2.astype(np.float32)
        assert np.array_equal(out[f"b2_{i}"].astype(np.float64), b2)  # float32 is lossless here
        if b3 is not None:
            out[f"b3_{i}"] = b3.astype(np.float32)
        if c is not None:
            out[f"cam_{i}"] = c.astype(np.float32)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(batches)} batches, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    for name, cfg in CASES.items():
        build(name, cfg)
