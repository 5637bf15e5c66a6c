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
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


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
