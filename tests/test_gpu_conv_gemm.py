"""GPU: operator-level parity of the tcgen05 conv GEMM (through the C ABI) against an fp64 torch
expectation computed from the same bf16-rounded operands.

Tolerances: with identical bf16 operands the only differences are fp32 accumulation order and the
output rounding: fp32 outputs must agree to 2e-5 of the output scale, bf16 outputs to one bf16 ulp
(2^-8 relative) of the expectation; bf16x3 (split operands) to 3e-5 of the fp32-operand product."""
import pytest
import torch

from gpu_utils import conv_gemm, expected_conv, pack_weight, planes_value, split_planes

pytestmark = pytest.mark.gpu


def _scale_err(got, exp):
    return float((got.double() - exp).abs().max() / exp.abs().max().clamp_min(1e-30))


def _mk(rows, ld, planes, dev, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.rand(rows, ld, generator=g) * 2 - 1).to(dev)
    return split_planes(a, planes)


@pytest.mark.parametrize("M,K,n_pad,ncols", [(300, 128, 64, 51), (1000, 1024, 256, 256),
                                             (129, 64, 128, 128), (64, 192, 64, 3)])
def test_flat_gemm_fp32_out(cuda_device, M, K, n_pad, ncols):
    dev = cuda_device
    a = _mk(M, K, 1, dev, 1)
    g = torch.Generator().manual_seed(2)
    w = ((torch.rand(ncols, K, 1, generator=g) * 2 - 1) / K ** 0.5).to(dev)
    wp = pack_weight(w, n_pad, K, 1)
    _, out = conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, out_f32_cols=ncols)
    exp = expected_conv(planes_value(a).reshape(M, K), planes_value(wp), samples=1, a_rows=M, taps=1,
                        k_per_tap=K, per_sample_tiles=False, tap_row_step=0, tap_col_step=0,
                        out_rows=M)[:, :ncols]
    assert not torch.isnan(out).any()
    assert _scale_err(out, exp) < 2e-5


def test_persistent_many_tiles(cuda_device):
    """More tiles than SMs x accumulator stages: exercises stage / phase wrap-around."""
    dev = cuda_device
    M, K, n_pad = 128 * 700 + 17, 192, 64
    a = _mk(M, K, 1, dev, 3)
    g = torch.Generator().manual_seed(4)
    w = ((torch.rand(n_pad, K, 1, generator=g) * 2 - 1) / K ** 0.5).to(dev)
    wp = pack_weight(w, n_pad, K, 1)
    _, out = conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, out_f32_cols=n_pad)
    exp = planes_value(a).reshape(M, K) @ planes_value(wp)[0].T
    assert _scale_err(out, exp) < 2e-5


def test_strided_conv_affine_relu_bf16_out(cuda_device):
    """Optimized1f block conv: stride == width == 3, taps are column blocks of one K = 3C row."""
    dev = cuda_device
    C, M = 128, 500
    a = _mk(M, 3 * C, 1, dev, 5)
    g = torch.Generator().manual_seed(6)
    w = ((torch.rand(C, C, 3, generator=g) * 2 - 1) / (3 * C) ** 0.5).to(dev)
    wp = pack_weight(w, C, C, 1)
    scale = (torch.rand(C, generator=g) + 0.5).to(dev)
    shift = (torch.randn(C, generator=g) * 0.1).to(dev)
    out, _ = conv_gemm(a, 1, M, 3 * C, wp, 3, C, C, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=C, out_rows=M, scale=scale, shift=shift, relu=True)
    acc = expected_conv(planes_value(a).reshape(M, 3 * C), planes_value(wp), samples=1, a_rows=M,
                        taps=3, k_per_tap=C, per_sample_tiles=False, tap_row_step=0, tap_col_step=C,
                        out_rows=M)
    exp = torch.relu(acc * scale.double() + shift.double())
    got = out[0].double()
    assert torch.all((got - exp).abs() <= exp.abs() * 2 ** -8 + 1e-6)


@pytest.mark.parametrize("dilation,L,samples", [(1, 40, 3), (9, 200, 2), (27, 300, 1)])
def test_dilated_conv_with_residual(cuda_device, dilation, L, samples):
    """TemporalModel block: 3 taps `dilation` frames apart, per-sample tiles (ragged last tile),
    then the residual slice-add of model.py:130-135 in the epilogue."""
    dev = cuda_device
    C = 128
    Lout = L - 2 * dilation
    a = _mk(samples * L, C, 1, dev, 7)
    g = torch.Generator().manual_seed(8)
    w = ((torch.rand(C, C, 3, generator=g) * 2 - 1) / (3 * C) ** 0.5).to(dev)
    wp = pack_weight(w, C, C, 1)
    scale = (torch.rand(C, generator=g) + 0.5).to(dev)
    shift = (torch.randn(C, generator=g) * 0.1).to(dev)
    out, _ = conv_gemm(a, samples, L, C, wp, 3, C, C, per_sample_tiles=True, tap_row_step=dilation,
                       tap_col_step=0, out_rows=Lout, scale=scale, shift=shift, relu=True, res=a,
                       res_rows_per_sample=L, res_row_step=1, res_row_off=dilation)
    av = planes_value(a).reshape(samples * L, C)
    acc = expected_conv(av, planes_value(wp), samples=samples, a_rows=L, taps=3, k_per_tap=C,
                        per_sample_tiles=True, tap_row_step=dilation, tap_col_step=0, out_rows=Lout)
    res = av.reshape(samples, L, C)[:, dilation:dilation + Lout].reshape(samples * Lout, C)
    exp = torch.relu(acc * scale.double() + shift.double()) + res
    got = out[0].double()
    assert not torch.isnan(got).any()
    assert torch.all((got - exp).abs() <= exp.abs() * 2 ** -8 + 1e-6)


def test_flat_residual_with_sample_div(cuda_device):
    """1x1 conv over rows flattened across samples, residual row = sample*L_in + t + off."""
    dev = cuda_device
    C, samples, Lin, Lout, off = 64, 5, 50, 44, 3
    a = _mk(samples * Lout, C, 1, dev, 9)
    r = _mk(samples * Lin, C, 1, dev, 10)
    g = torch.Generator().manual_seed(11)
    w = ((torch.rand(C, C, 1, generator=g) * 2 - 1) / C ** 0.5).to(dev)
    wp = pack_weight(w, C, C, 1)
    out, _ = conv_gemm(a, 1, samples * Lout, C, wp, 1, C, C, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=samples * Lout, res=r, res_rows_per_sample=Lin,
                       res_row_step=1, res_row_off=off, res_sample_div=Lout)
    acc = planes_value(a).reshape(-1, C) @ planes_value(wp)[0].T
    res = planes_value(r).reshape(samples, Lin, C)[:, off:off + Lout].reshape(-1, C)
    exp = acc + res
    assert torch.all((out[0].double() - exp).abs() <= exp.abs() * 2 ** -8 + 1e-6)


def test_bf16x3_is_fp32_faithful(cuda_device):
    dev = cuda_device
    M, K, n_pad = 700, 1024, 256
    g = torch.Generator().manual_seed(12)
    a32 = (torch.rand(M, K, generator=g) * 2 - 1).to(dev)
    w32 = ((torch.rand(n_pad, K, 1, generator=g) * 2 - 1) / K ** 0.5).to(dev)
    a = split_planes(a32, 2)
    wp = pack_weight(w32, n_pad, K, 2)
    out, _ = conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, precision=1, out_planes=2)
    exp = a32.double() @ w32[:, :, 0].double().T
    got = planes_value(out)
    assert _scale_err(got, exp) < 3e-5


def test_stats_accumulation(cuda_device):
    dev = cuda_device
    M, K, n_pad = 1000, 128, 256
    a = _mk(M, K, 1, dev, 13)
    g = torch.Generator().manual_seed(14)
    w = ((torch.rand(n_pad, K, 1, generator=g) * 2 - 1) / K ** 0.5).to(dev)
    wp = pack_weight(w, n_pad, K, 1)
    # per-slab partials: [4 * row tiles][2][n_pad], slab s = rows 32*s .. 32*s+31
    slabs = 4 * ((M + 127) // 128)
    stats = torch.full((slabs, 2, n_pad), float("nan"), dtype=torch.float32, device=dev)
    out, _ = conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, stats=stats)
    acc = planes_value(a).reshape(M, K) @ planes_value(wp)[0].T
    assert not torch.isnan(stats).any()                      # every slab entry is written
    pad = torch.zeros(slabs * 32 - M, n_pad, dtype=acc.dtype, device=dev)
    per_slab = torch.cat([acc, pad]).reshape(slabs, 32, n_pad)
    assert _scale_err(stats[:, 0], per_slab.sum(1)) < 1e-5
    assert _scale_err(stats[:, 1], (per_slab * per_slab).sum(1)) < 1e-5
    # bit-identical on a second launch (plain stores, no atomics)
    stats2 = torch.zeros_like(stats)
    conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
              tap_col_step=0, out_rows=M, stats=stats2)
    assert torch.equal(stats, stats2)


# ---- fp16 operand format (VP3D_PRECISION_FP16 = 3): the eval default
def _mk16(rows, ld, dev, seed, amp=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = ((torch.rand(rows, ld, generator=g) * 2 - 1) * amp).to(dev)
    return a.to(torch.float16).unsqueeze(0).contiguous()


def _pack16(w, n_pad, k_pad):
    co, ci, k = w.shape
    buf = torch.zeros(k, n_pad, k_pad, dtype=torch.float32, device=w.device)
    buf[:, :co, :ci] = w.permute(2, 0, 1)
    return buf.to(torch.float16).unsqueeze(0).contiguous()


@pytest.mark.parametrize("M,K,n_pad,ncols", [(300, 128, 64, 51), (1000, 1024, 256, 256)])
def test_fp16_flat_gemm_fp32_out(cuda_device, M, K, n_pad, ncols):
    dev = cuda_device
    a = _mk16(M, K, dev, 21)
    g = torch.Generator().manual_seed(22)
    w = ((torch.rand(ncols, K, 1, generator=g) * 2 - 1) / K ** 0.5).to(dev)
    wp = _pack16(w, n_pad, K)
    _, out = conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, out_f32_cols=ncols, precision=3)
    exp = (a[0].double() @ wp[0, 0].double().T)[:, :ncols]
    assert not torch.isnan(out).any()
    assert _scale_err(out, exp) < 2e-5


def test_fp16_strided_conv_residual_fp16_out(cuda_device):
    """1x1 conv + affine + ReLU + TMA-loaded residual (the block tail of the eval cone schedule) in
    fp16 storage: result within one fp16 ulp (2^-11 relative) of the fp64 expectation."""
    dev = cuda_device
    C, M = 256, 700
    a = _mk16(M, C, dev, 23)
    r = _mk16(3 * M, C, dev, 24, amp=4.0)
    g = torch.Generator().manual_seed(25)
    w = ((torch.rand(C, C, 1, generator=g) * 2 - 1) / C ** 0.5).to(dev)
    wp = _pack16(w, C, C)
    scale = (torch.rand(C, generator=g) + 0.5).to(dev)
    shift = (torch.randn(C, generator=g) * 0.1).to(dev)
    out, _ = conv_gemm(a, 1, M, C, wp, 1, C, C, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, scale=scale, shift=shift, relu=True, res=r,
                       res_rows_per_sample=0, res_row_step=1, res_row_off=M, precision=3)
    acc = a[0].double() @ wp[0, 0].double().T
    exp = torch.relu(acc * scale.double() + shift.double()) + r[0, M:2 * M].double()
    got = out[0].double()
    assert out.dtype == torch.float16 and not torch.isnan(got).any()
    assert torch.all((got - exp).abs() <= exp.abs() * 2 ** -11 + 1e-6)


def test_fp16_store_saturates(cuda_device):
    """An activation beyond the fp16 range clamps to +-65504 instead of becoming inf."""
    dev = cuda_device
    M, K, n_pad = 128, 64, 64
    a = torch.full((1, M, K), 60.0, dtype=torch.float16, device=dev)
    w = torch.full((n_pad, K, 1), 30.0, device=dev)
    w[1::2] *= -1
    wp = _pack16(w, n_pad, K)
    out, _ = conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, precision=3)
    got = out[0].float()
    assert torch.isfinite(got).all()
    assert torch.all(got[:, 0::2] == 65504.0) and torch.all(got[:, 1::2] == -65504.0)
