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


# ---- 256-wide launches with many row tiles: these run on CTA pairs (tcgen05.mma.cta_group::2)
# unless VP3D_PAIR=0; the expectations are the same either way.
def test_wide_flat_gemm_many_tiles_odd_count(cuda_device):
    """151 row tiles (odd: the last pair has one out-of-range member), ragged last tile, fp32 out."""
    dev = cuda_device
    M, K, n_pad = 128 * 150 + 37, 192, 512
    a = _mk(M, K, 1, dev, 31)
    g = torch.Generator().manual_seed(32)
    w = ((torch.rand(n_pad, K, 1, generator=g) * 2 - 1) / K ** 0.5).to(dev)
    wp = pack_weight(w, n_pad, K, 1)
    _, out = conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, out_f32_cols=n_pad)
    exp = planes_value(a).reshape(M, K) @ planes_value(wp)[0].T
    assert not torch.isnan(out).any()
    assert _scale_err(out, exp) < 2e-5


@pytest.mark.parametrize("precision", [0, 3])
def test_wide_strided_block_tail_with_tma_residual(cuda_device, precision):
    """The eval cone schedule's block tail at 256-wide tiles: 3-tap conv over row regions, then a
    1x1 conv + affine + ReLU + TMA-loaded residual, bf16 and fp16 storage; K = 1024 keeps the
    operand pipeline wrapping many times per tile."""
    dev = cuda_device
    C, R = 256, 128 * 101 + 5            # R output rows, 3R input rows (tap-major regions)
    dt = torch.float16 if precision == 3 else torch.bfloat16
    g = torch.Generator().manual_seed(33)
    x = ((torch.rand(3 * R, C, generator=g) * 2 - 1)).to(dev).to(dt).unsqueeze(0).contiguous()
    w3 = ((torch.rand(C, C, 3, generator=g) * 2 - 1) / (3 * C) ** 0.5).to(dev)
    w1 = ((torch.rand(C, C, 1, generator=g) * 2 - 1) / C ** 0.5).to(dev)

    def pk(w):
        co, ci, k = w.shape
        return w.permute(2, 0, 1).contiguous().to(dt).unsqueeze(0).contiguous()
    scale = (torch.rand(C, generator=g) + 0.5).to(dev)
    shift = (torch.randn(C, generator=g) * 0.1).to(dev)
    h, _ = conv_gemm(x, 1, 3 * R, C, pk(w3), 3, C, C, per_sample_tiles=False, tap_row_step=R,
                     tap_col_step=0, out_rows=R, scale=scale, shift=shift, relu=True,
                     precision=precision)
    xv = x[0].double()
    acc = sum(xv[k * R:(k + 1) * R] @ pk(w3)[0, k].double().T for k in range(3))
    h_exp = torch.relu(acc * scale.double() + shift.double())
    ulp = 2 ** -11 if precision == 3 else 2 ** -8
    assert torch.all((h[0].double() - h_exp).abs() <= h_exp.abs() * ulp + 1e-6)
    y, _ = conv_gemm(h, 1, R, C, pk(w1), 1, C, C, per_sample_tiles=False, tap_row_step=0,
                     tap_col_step=0, out_rows=R, scale=scale, shift=shift, relu=True, res=x,
                     res_rows_per_sample=0, res_row_step=1, res_row_off=R, precision=precision)
    acc = h[0].double() @ pk(w1)[0, 0].double().T
    y_exp = torch.relu(acc * scale.double() + shift.double()) + xv[R:2 * R]
    assert not torch.isnan(y[0].double()).any()
    assert torch.all((y[0].double() - y_exp).abs() <= y_exp.abs() * ulp + 1e-6)


def test_wide_dilated_tiles_and_stats(cuda_device):
    """Per-sample tiles (dilated layout, 3 ragged tiles per sample, odd tile count) at 256-wide
    tiles with the training epilogue: raw bf16 store + per-slab batch statistics."""
    dev = cuda_device
    C, samples, L, dil = 256, 51, 300 + 18, 9
    Lout = L - 2 * dil
    a = _mk(samples * L, C, 1, dev, 34)
    g = torch.Generator().manual_seed(35)
    w = ((torch.rand(C, C, 3, generator=g) * 2 - 1) / (3 * C) ** 0.5).to(dev)
    wp = pack_weight(w, C, C, 1)
    tiles = samples * ((Lout + 127) // 128)
    stats = torch.full((4 * tiles, 2, C), float("nan"), dtype=torch.float32, device=dev)
    out, _ = conv_gemm(a, samples, L, C, wp, 3, C, C, per_sample_tiles=True, tap_row_step=dil,
                       tap_col_step=0, out_rows=Lout, stats=stats)
    av = planes_value(a).reshape(samples * L, C)
    acc = expected_conv(av, planes_value(wp), samples=samples, a_rows=L, taps=3, k_per_tap=C,
                        per_sample_tiles=True, tap_row_step=dil, tap_col_step=0, out_rows=Lout)
    assert torch.all((out[0].double() - acc).abs() <= acc.abs() * 2 ** -8 + 1e-6)
    assert not torch.isnan(stats).any()
    assert _scale_err(stats[:, 0].sum(0), acc.sum(0)) < 1e-4
    assert _scale_err(stats[:, 1].sum(0), (acc * acc).sum(0)) < 1e-4


def test_wide_bf16x3_two_planes(cuda_device):
    dev = cuda_device
    M, K, n_pad = 128 * 80 + 9, 256, 256
    g = torch.Generator().manual_seed(36)
    a32 = (torch.rand(M, K, generator=g) * 2 - 1).to(dev)
    w32 = ((torch.rand(n_pad, K, 1, generator=g) * 2 - 1) / K ** 0.5).to(dev)
    a = split_planes(a32, 2)
    wp = pack_weight(w32, n_pad, K, 2)
    out, _ = conv_gemm(a, 1, M, K, wp, 1, K, n_pad, per_sample_tiles=False, tap_row_step=0,
                       tap_col_step=0, out_rows=M, precision=1, out_planes=2)
    exp = a32.double() @ w32[:, :, 0].double().T
    assert _scale_err(planes_value(out), exp) < 3e-5
