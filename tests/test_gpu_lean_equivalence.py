"""GPU: the compile-time "lean" inference epilogue of conv_gemm_kernel (paired FMAs, ReLU applied to
the packed 16-bit pairs, result staged in place in the residual landing tile, weight tiles primed
ahead of the dependency wait) must reproduce the general epilogue BIT FOR BIT: fma.rn.f32x2 rounds
each lane like fmaf, max(·, 0) commutes with the 16-bit rounding, and the in-place staging moves no
arithmetic.  The switch (VP3D_LEAN) is read once per process, so both variants run in children.
Reference semantics: common/model.py:63-77, 126-138, 187-197 (eval forward of both classes)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CHILD = r"""
import sys, numpy as np, torch
sys.path.insert(0, %(root)r)
import videopose3d_b200 as vp
out = {}
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(11)
cases = [
    ("tm_cone_fp16", vp.TemporalModel, [3, 3, 3], 128, 96, 27, "fp16"),      # strided eval schedule
    ("tm_cone_bf16", vp.TemporalModel, [3, 3, 3], 128, 96, 27, "bf16"),
    ("tm_seq_fp16", vp.TemporalModel, [3, 3, 3], 128, 3, 300, "fp16"),       # dilated schedule
    ("opt_fp16", vp.TemporalModelOptimized1f, [3, 3, 3], 256, 640, 27, "fp16"),
    ("tm_c100_fp16", vp.TemporalModel, [3, 5], 100, 64, 15, "fp16"),         # padded channels
]
for name, cls, arc, ch, n, t, prec in cases:
    torch.manual_seed(3)
    m = cls(17, 2, 17, filter_widths=arc, channels=ch).to(dev).eval().set_precision(prec)
    with torch.no_grad():
        for bn in [m.expand_bn] + list(m.layers_bn):     # non-trivial running statistics
            bn.running_mean.uniform_(-0.2, 0.2, generator=None)
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.3, 0.3)
        x = (torch.rand(n, t, 17, 2, generator=g) * 2 - 1).to(dev)
        out[name] = m(x).float().cpu().numpy()
np.savez(sys.argv[1], **out)
"""


def _run(lean, path):
    env = dict(os.environ, VP3D_LEAN=lean)
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}, path], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(path)


def test_lean_epilogue_matches_general_epilogue_bitwise(tmp_path):
    a = _run("1", str(tmp_path / "lean.npz"))
    b = _run("0", str(tmp_path / "general.npz"))
    assert set(a.files) == set(b.files) and len(a.files) == 5
    for k in a.files:
        assert a[k].shape == b[k].shape
        assert np.isfinite(a[k]).all()
        assert np.array_equal(a[k], b[k]), (k, float(np.abs(a[k] - b[k]).max()))
