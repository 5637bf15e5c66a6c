"""CPU: pin the oracle (both restatements) against the golden vectors produced by the real
reference (tests/golden/make_golden.py).  Tolerances: the NumPy float64 restatement must agree with
the reference's float32 CPU output to float32 round-off (<= 2e-5 of the output scale); the
torch.nn.functional restatement (float32) likewise."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import temporal_model_oracle as orc

SMALL = [n for n in golden_names() if "c1024" not in n]
LARGE = [n for n in golden_names() if "c1024" in n]


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.mark.parametrize("name", SMALL + LARGE)
def test_numpy_restatement_matches_reference(name):
    meta, sd, x, y_ref, new = load_golden(name)
    if meta["C"] == 1024 and meta["N"] * meta["T"] > 2000:
        x, y_ref = x[:2], y_ref[:2]  # keep the CPU suite fast; batch rows are independent in eval
    strided = meta["cls"] == "TemporalModelOptimized1f"
    train = bool(meta.get("train"))
    res = orc.forward_numpy(sd, x.numpy(), meta["fw"], causal=meta["causal"], dense=meta["dense"],
                            strided=strided, training=train, momentum=meta.get("momentum", 0.1),
                            return_new_stats=train)
    y = res[0] if train else res
    assert y.shape == y_ref.shape
    assert _rel(y, y_ref) < 2e-5
    if train:
        for k, v in new.items():
            if k == "gy" or k.startswith("grad/"):
                continue
            if k.endswith("num_batches_tracked"):
                assert int(res[1][k]) == int(v)
            else:
                assert _rel(res[1][k], v) < 2e-5, k


@pytest.mark.parametrize("name", SMALL + LARGE)
def test_torch_restatement_matches_reference(name):
    meta, sd, x, y_ref, new = load_golden(name)
    if meta["C"] == 1024 and meta["N"] * meta["T"] > 2000:
        x, y_ref = x[:2], y_ref[:2]
    strided = meta["cls"] == "TemporalModelOptimized1f"
    train = bool(meta.get("train"))
    sd = {k: v.clone() for k, v in sd.items()}
    if train:
        for k, v in sd.items():
            if v.dtype.is_floating_point and "running" not in k:
                v.requires_grad_(True)
    with torch.set_grad_enabled(train):
        y = orc.forward_torch(sd, x, meta["fw"], causal=meta["causal"], dense=meta["dense"],
                              strided=strided, training=train, momentum=meta.get("momentum", 0.1),
                              update_stats=train)
    assert tuple(y.shape) == y_ref.shape
    assert _rel(y.detach().numpy(), y_ref) < 2e-5
    if train:
        (y * torch.from_numpy(new["gy"])).sum().backward()
        for k, v in new.items():
            if k == "gy" or k.endswith("num_batches_tracked"):
                continue
            if k.startswith("grad/"):
                # the oracle's autograd gradients == the reference's (both fp32 on CPU)
                assert _rel(sd[k[5:]].grad.numpy(), v) < 1e-4, k
            else:
                assert _rel(sd[k].detach().numpy(), v) < 2e-5, k


def test_eval_cone_equals_dilated():
    """The property the strided eval schedule relies on: with running statistics,
    TemporalModel on one receptive field == TemporalModelOptimized1f, causal or not."""
    for name in ("tm_333_c64_rf", "tm_333_c64_rf_causal"):
        meta, sd, x, y_ref, _ = load_golden(name)
        y = orc.forward_numpy(sd, x.numpy(), meta["fw"], causal=meta["causal"], strided=True)
        assert _rel(y, y_ref) < 2e-5


def test_arch_bookkeeping():
    a = orc.arch([3, 3, 3, 3, 3])
    assert a["receptive_field"] == 243 and a["pad"] == [1, 3, 9, 27, 81]
    assert orc.arch([3, 3, 3], causal=True)["shift"] == [1, 3, 9]
    assert orc.arch([3, 3, 3], causal=True, strided=True)["shift"] == [1, 1, 1]
    with pytest.raises(AssertionError):
        orc.arch([3, 4])


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("opt") and "train" in n])
def test_train_emulation_matches_reference(name):
    """oracle/train_emulation.py (analytic backward, float64, no rounding) == the reference's
    autograd gradients, updated running statistics and output."""
    from oracle import train_emulation as emu
    meta, sd, x, y_ref, new = load_golden(name)
    out = emu.train_step(sd, x, torch.from_numpy(new["gy"]), meta["fw"], causal=meta["causal"],
                         planes=0, momentum=meta["momentum"])
    assert out["min_abs_preact"] >= 2e-4, "fixture sits on a ReLU kink"
    assert _rel(out["y"].numpy(), y_ref) < 2e-5
    for k, v in out["grads"].items():
        assert _rel(v.numpy(), new["grad/" + k]) < 1e-4, k
    for k, v in out["new_stats"].items():
        assert _rel(v.numpy(), new[k]) < 2e-5, k
    # split-bf16 rounding points keep the step fp32-faithful on a well-conditioned fixture
    out2 = emu.train_step(sd, x, torch.from_numpy(new["gy"]), meta["fw"], causal=meta["causal"],
                          planes=2, momentum=meta["momentum"])
    for k, v in out2["grads"].items():
        assert _rel(v.numpy(), new["grad/" + k]) < 1e-3, k
