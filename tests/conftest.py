import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    """Model fixtures (tests/golden/make_golden.py)."""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    # gen_*: generator fixtures; semi_* / big_*: tests/golden/make_semi_golden.py (own loaders)
    return [n for n in names if not n.startswith(("gen_", "semi_", "big_"))]


def generator_golden_names():
    """Generator fixtures (tests/golden/make_generator_golden.py)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "gen_*.npz")))


def load_generator_golden(name):
    """Return (cfg, cameras|None, poses_3d|None, poses_2d, batches) where batches is a list of
    (cam|None, b3|None, b2) float32 arrays exactly as the reference generator yielded them."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = json.loads(str(z["config"]))
    n = len(cfg["lengths"])
    p2 = [z[f"p2_{i}"] for i in range(n)]
    p3 = [z[f"p3_{i}"] for i in range(n)] if cfg["use_3d"] else None
    cams = list(z["cams"]) if cfg["use_cam"] else None
    batches = []
    for i in range(int(z["n_batches"])):
        batches.append((z[f"cam_{i}"] if f"cam_{i}" in z.files else None,
                        z[f"b3_{i}"] if f"b3_{i}" in z.files else None, z[f"b2_{i}"]))
    return cfg, cams, p3, p2, batches


def generator_kwargs(cfg):
    """Constructor keywords shared by the reference generators, the oracle and the device classes."""
    kw = dict(pad=cfg["pad"], causal_shift=cfg["causal_shift"], augment=cfg["augment"],
              kps_left=cfg["left"], kps_right=cfg["right"], joints_left=cfg["left"],
              joints_right=cfg["right"])
    if cfg["kind"] == "chunked":
        kw.update(shuffle=cfg["shuffle"], random_seed=cfg["random_seed"],
                  endless=cfg.get("endless", False))
    return kw


def load_golden(name):
    """Return (meta, state_dict(torch), x(torch), y(numpy), new_stats(dict) ) for a fixture."""
    import torch
    from oracle import temporal_model_oracle as orc
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    if "x" in z.files:
        sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        x = torch.from_numpy(z["x"])
    else:
        sd = orc.make_state_dict(meta["J"], meta["F"], meta["Jout"], meta["fw"], meta["C"],
                                 dense=meta["dense"], seed=meta["seed"])
        x = orc.make_input(meta["N"], meta["T"], meta["J"], meta["F"], seed=meta["seed"] + 1)
    new = {k[4:]: z[k] for k in z.files if k.startswith("new/")}
    if "gy" in z.files:
        new["gy"] = z["gy"]
        new.update({k: z[k] for k in z.files if k.startswith("grad/")})
    return meta, sd, x, z["y"], new


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
