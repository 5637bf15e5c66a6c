"""GPU: model-level parity of the CUDA path (nn.Module -> C ABI -> sm_100a kernels) against
(a) the golden vectors produced by the real reference and (b) the oracle on fresh seeded inputs.

Gates:
  G1  default (fp16 operands, fp32 accumulate) and bf16x3 (split-bf16) modes:
      max|new - ref| / max|ref| <= 1e-3 on every golden (north_star's tolerance; measured ~4e-4 /
      ~1e-5)
  G2  |MPJPE(new, y) - MPJPE(ref, y)| <= 0.1 mm on synthetic targets y = ref + N(0, 0.03^2) m at
      BASELINE configs[1] size, plus mpjpe(new, ref) reported
  secondary modes: `mixed` <= 2e-3 (C = 1024) / 3e-3, pure `bf16` <= 3e-2 (documented, not default)
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import temporal_model_oracle as orc
import videopose3d_b200 as vp

pytestmark = pytest.mark.gpu


def _build(meta, sd, dev, precision):
    kw = dict(filter_widths=meta["fw"], causal=meta["causal"], dropout=0.0, channels=meta["C"])
    if meta["cls"] == "TemporalModel":
        m = vp.TemporalModel(meta["J"], meta["F"], meta["Jout"], dense=meta["dense"], **kw)
    else:
        m = vp.TemporalModelOptimized1f(meta["J"], meta["F"], meta["Jout"], **kw)
    m.load_state_dict(sd)
    return m.to(dev).eval().set_precision(precision)


def _rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())


EVAL_CASES = [n for n in golden_names() if "train" not in n]


@pytest.mark.parametrize("name", EVAL_CASES)
def test_golden_eval_fp32_faithful(cuda_device, name):
    meta, sd, x, y_ref, _ = load_golden(name)
    m = _build(meta, sd, cuda_device, "bf16x3")
    with torch.no_grad():
        y = m(x.to(cuda_device))
    assert tuple(y.shape) == y_ref.shape and y.dtype == torch.float32
    assert _rel(y.cpu().numpy(), y_ref) <= 1e-3


@pytest.mark.parametrize("name", EVAL_CASES)
def test_golden_eval_default_fp16(cuda_device, name):
    """The DEFAULT (and benchmarked) precision mode -- fp16 operands / activations, fp32
    accumulate -- holds north_star's 1e-3 on every golden produced by the real reference."""
    meta, sd, x, y_ref, _ = load_golden(name)
    kw = dict(filter_widths=meta["fw"], causal=meta["causal"], dropout=0.0, channels=meta["C"])
    if meta["cls"] == "TemporalModel":
        m = vp.TemporalModel(meta["J"], meta["F"], meta["Jout"], dense=meta["dense"], **kw)
    else:
        m = vp.TemporalModelOptimized1f(meta["J"], meta["F"], meta["Jout"], **kw)
    m.load_state_dict(sd)
    m = m.to(cuda_device).eval()
    assert m.precision == "fp16"  # nothing selected: this is what run.py gets
    with torch.no_grad():
        y = m(x.to(cuda_device)).cpu()
    assert tuple(y.shape) == y_ref.shape and y.dtype == torch.float32
    assert _rel(y.numpy(), y_ref) <= 1e-3


@pytest.mark.parametrize("name", EVAL_CASES)
def test_golden_eval_mixed(cuda_device, name):
    """Secondary mode `mixed` (bf16 on the FLOP-dominant blocks, split-bf16 elsewhere, exact
    residual stream): <= 3e-3 of the output scale on every golden (<= 2e-3 on the C = 1024 ones;
    measured 0.7e-3 .. 1.5e-3 -- which is why it is no longer the default)."""
    meta, sd, x, y_ref, _ = load_golden(name)
    m = _build(meta, sd, cuda_device, "mixed")
    with torch.no_grad():
        y = m(x.to(cuda_device)).cpu()
    assert _rel(y.numpy(), y_ref) <= (2e-3 if meta["C"] == 1024 else 3e-3)


@pytest.mark.parametrize("name", EVAL_CASES)
def test_golden_eval_bf16(cuda_device, name):
    """pure-bf16 mode against the reference goldens: bounded relative error (bf16 operands through
    2B+2 layers) and mpjpe(new, ref) below 1% of the output scale.  The 0.1 mm MPJPE-shift gate
    needs thousands of joints to be a statistic rather than noise: see test_bf16_mpjpe_gate."""
    meta, sd, x, y_ref, _ = load_golden(name)
    m = _build(meta, sd, cuda_device, "bf16")
    with torch.no_grad():
        y = m(x.to(cuda_device)).cpu()
    assert _rel(y.numpy(), y_ref) <= 3e-2
    ref = torch.from_numpy(y_ref)
    scale = float(torch.norm(ref, dim=-1).mean())
    assert float(orc.mpjpe(y, ref)) <= 1e-2 * scale


def test_default_mpjpe_gate(cuda_device):
    """G2 at BASELINE configs[1] size (N = 1024 windows -> 17408 joints): MPJPE of the default
    (fp16) path
    against synthetic targets y = ref + N(0, 30 mm) differs from the reference's MPJPE by <= 0.1 mm
    (outputs read as metres).  `ref` here is the fp32-faithful CUDA path, itself pinned to the
    reference goldens at <= 1e-3 by test_golden_eval_fp32_faithful."""
    meta, sd, x8, y_ref, _ = load_golden("cfg2_tm_33333_c1024")
    x = orc.make_input(1024, 243, seed=78)
    x[:8] = x8
    xg = x.to(cuda_device)
    m = _build(meta, sd, cuda_device, "bf16x3")
    with torch.no_grad():
        ref = m(xg).cpu()
        y = m.set_precision("fp16")(xg).cpu()
        y_pure = m.set_precision("bf16")(xg).cpu()
        y_mixed = m.set_precision("mixed")(xg).cpu()
    assert _rel(ref[:8].numpy(), y_ref) <= 1e-3
    assert _rel(y[:8].numpy(), y_ref) <= 1e-3
    # whole batch against the fp32-faithful CUDA path (itself <= 1e-3, measured ~1e-5, of the goldens)
    full = float((y - ref).abs().max() / ref.abs().max())
    print(f"fp16 vs bf16x3 over 1024 windows: rel {full:.2e}; "
          f"mixed {float((y_mixed - ref).abs().max() / ref.abs().max()):.2e}; "
          f"bf16 {float((y_pure - ref).abs().max() / ref.abs().max()):.2e}")
    assert full <= 1e-3
    print(f"pure bf16: mpjpe(bf16,ref)={float(orc.mpjpe(y_pure, ref)) * 1000:.3f} mm, "
          f"mixed: {float(orc.mpjpe(y_mixed, ref)) * 1000:.3f} mm")
    g = torch.Generator().manual_seed(5)
    target = ref + torch.randn(ref.shape, generator=g) * 0.03
    target[:, :, 0] = ref[:, :, 0]
    e_new, e_ref = float(orc.mpjpe(y, target)) * 1000, float(orc.mpjpe(ref, target)) * 1000
    direct = float(orc.mpjpe(y, ref)) * 1000
    print(f"MPJPE(ref,y)={e_ref:.3f} mm  MPJPE(fp16,y)={e_new:.3f} mm  mpjpe(fp16,ref)={direct:.3f} mm "
          f"output scale {float(torch.norm(ref, dim=-1).mean()):.3f}")
    assert abs(e_new - e_ref) <= 0.1


def test_output_is_fresh_writable_tensor(cuda_device):
    """run.py:677-679 mutates the returned tensor in place."""
    meta, sd, x, y_ref, _ = load_golden("tm_333_c64")
    m = _build(meta, sd, cuda_device, "bf16x3")
    with torch.no_grad():
        y1 = m(x.to(cuda_device))
        y1[:, :, :, 0] *= -1
        y2 = m(x.to(cuda_device))
    assert y1.data_ptr() != y2.data_ptr()
    assert _rel(y2.cpu().numpy(), y_ref) <= 1e-3


def test_weights_are_repacked_after_update(cuda_device):
    meta, sd, x, y_ref, _ = load_golden("opt_333_c64")
    m = _build(meta, sd, cuda_device, "bf16x3")
    xg = x.to(cuda_device)
    with torch.no_grad():
        y0 = m(xg).cpu().numpy()
        m.shrink.bias.add_(1.0)
        m.layers_bn[1].running_mean.mul_(0.5)
        m.layers_conv[0].weight.mul_(1.1)
        y1 = m(xg).cpu().numpy()
    sd2 = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    y_o = orc.forward_numpy(sd2, x.numpy(), meta["fw"], causal=meta["causal"], strided=True)
    assert _rel(y0, y_ref) <= 1e-3
    assert _rel(y1, y_o) <= 1e-3
    assert _rel(y1, y_ref) > 1e-2


def test_cfg2_full_batch_properties(cuda_device):
    """BASELINE configs[1] at full size (N = 1024, T = 243): size-independent properties — batch
    rows are independent in eval mode, so any sub-batch must reproduce the full-batch rows exactly,
    and rows shared with the committed N = 8 golden must match it."""
    meta, sd, x8, y_ref, _ = load_golden("cfg2_tm_33333_c1024")
    m = _build(meta, sd, cuda_device, "bf16")
    x = orc.make_input(1024, 243, seed=77)
    x[:8] = x8
    xg = x.to(cuda_device)
    with torch.no_grad():
        y = m(xg)
        y_sub = m(xg[500:564].contiguous())
    assert tuple(y.shape) == (1024, 1, 17, 3)
    assert torch.isfinite(y).all()
    assert torch.equal(y[500:564], y_sub)
    assert _rel(y[:8].cpu().numpy(), y_ref) <= 3e-2


def test_dilated_and_cone_schedules_agree(cuda_device):
    """T == RF uses the strided (cone) schedule; appending one frame forces the dilated schedule,
    whose first output frame sees the same receptive field."""
    meta, sd, x, y_ref, _ = load_golden("tm_333_c64_rf")
    m = _build(meta, sd, cuda_device, "bf16x3")
    xg = x.to(cuda_device)
    x_long = torch.cat([xg, xg[:, -1:]], dim=1).contiguous()
    with torch.no_grad():
        y_cone = m(xg)
        y_dil = m(x_long)[:, :1]
    assert float((y_cone - y_dil).abs().max() / y_cone.abs().max()) <= 1e-4


def test_forward_host_matches_device(cuda_device):
    meta, sd, x, y_ref, _ = load_golden("tm_333_c64")
    m = _build(meta, sd, cuda_device, "bf16x3")
    y_h = m.forward_host(x.pin_memory())
    assert not y_h.is_cuda
    assert _rel(y_h.numpy(), y_ref) <= 1e-3


def test_pipelined_host_api_matches_device(cuda_device):
    """Two-slot submit / wait: results of interleaved batches land in the right buffers."""
    meta, sd, x, y_ref, _ = load_golden("tm_333_c64")
    m = _build(meta, sd, cuda_device, "bf16x3")
    xs = [x.pin_memory(), (x * 0.5).pin_memory(), (-x).pin_memory()]
    with torch.no_grad():
        expect = [m(t.to(cuda_device)).cpu() for t in xs]
    outs = [torch.empty_like(expect[0]).pin_memory() for _ in xs]
    m.forward_host_submit(xs[0], outs[0], 0)
    m.forward_host_submit(xs[1], outs[1], 1)
    with pytest.raises(RuntimeError):
        m.forward_host_submit(xs[2], outs[2], 1)      # slot still in flight
    m.forward_host_wait(0)
    m.forward_host_submit(xs[2], outs[2], 0)
    m.forward_host_wait(1)
    m.forward_host_wait(0)
    for o, e in zip(outs, expect):
        assert torch.equal(o, e)
    assert _rel(outs[0].numpy(), y_ref) <= 1e-3


def test_errors(cuda_device):
    meta, sd, x, _, _ = load_golden("tm_333_c64")
    m = _build(meta, sd, cuda_device, "bf16")
    with pytest.raises(AssertionError):
        m(torch.zeros(2, 30, 16, 2, device=cuda_device))
    with pytest.raises(ValueError):
        m(torch.zeros(2, 10, 17, 2, device=cuda_device))   # shorter than the receptive field
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 30, 17, 2))                        # CPU tensor: no fallback
    with pytest.raises(AssertionError):
        vp.TemporalModel(17, 2, 17, [3, 4, 3])


def test_strided_model_reproduces_reference_size_mismatch_error(cuda_device):
    """Optimized1f on a length whose residual slice and strided conv disagree (T = 33, arc 3,3,3:
    11 frames after expand -> conv gives 3, x[:, :, 1::3] gives 4): the reference raises a size
    mismatch in `res + x` (model.py:191-194); so does the C ABI, with the same message."""
    meta, sd, x, _, _ = load_golden("opt_333_c64")
    m = _build(meta, sd, cuda_device, "fp16")
    with pytest.raises((ValueError, RuntimeError), match="must match the size of tensor b"):
        m(torch.zeros(2, 33, 17, 2, device=cuda_device))
    y = m(torch.zeros(2, 30, 17, 2, device=cuda_device))   # 30 = 27 + ignored trailing frames: fine
    assert tuple(y.shape) == (2, 1, 17, 3)


def test_any_channel_count(cuda_device):
    """`channels` need not be a multiple of 64 (run.py -ch N): odd sizes against the oracle."""
    for C in (1, 17, 96, 129):
        sd = orc.make_state_dict(17, 2, 17, [3, 3], C, seed=40 + C)
        x = orc.make_input(5, 13, 17, 2, seed=3)
        y_ref = orc.forward_numpy(sd, x.numpy(), [3, 3])
        m = vp.TemporalModel(17, 2, 17, filter_widths=[3, 3], channels=C)
        m.load_state_dict(sd)
        m = m.to(cuda_device).eval()
        with torch.no_grad():
            y16 = m(x.to(cuda_device)).cpu().numpy()
            y3 = m.set_precision("bf16x3")(x.to(cuda_device)).cpu().numpy()
        assert _rel(y3, y_ref) <= 1e-4, C
        assert _rel(y16, y_ref) <= 1e-3, C
