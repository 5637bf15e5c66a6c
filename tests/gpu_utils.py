"""Helpers for the GPU tests: build operands for the op-level C-ABI entry (vp3d_conv_gemm) from
plain torch tensors and compute the fp64 expectation of the same fused op."""
import ctypes

import torch

from videopose3d_b200 import _capi


def split_planes(t, planes):
    """fp32 tensor -> bf16 [planes, ...] (hi, lo) exactly as the kernels split values."""
    hi = t.to(torch.bfloat16)
    if planes == 1:
        return hi.unsqueeze(0).contiguous()
    lo = (t - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo]).contiguous()


def planes_value(p):
    """bf16 [planes, ...] -> the fp64 value the kernel operand represents."""
    return p.double().sum(dim=0)


def pack_weight(w, n_pad, k_pad, planes):
    """torch Conv1d weight (Cout, Cin, K) fp32 -> bf16 [planes][K][n_pad][k_pad]."""
    co, ci, k = w.shape
    buf = torch.zeros(k, n_pad, k_pad, dtype=torch.float32, device=w.device)
    buf[:, :co, :ci] = w.permute(2, 0, 1)
    return split_planes(buf, planes)


def conv_gemm(a_planes_t, samples, a_rows, a_ld, w_planes_t, taps, k_per_tap, n_pad, *,
              per_sample_tiles, tap_row_step, tap_col_step, out_rows, precision=0, scale=None,
              shift=None, relu=False, res=None, res_rows_per_sample=0, res_row_step=1, res_row_off=0,
              res_sample_div=0, out_planes=1, out_f32_cols=None, stats=None):
    """Launch vp3d_conv_gemm; returns (out_bf16_planes or None, out_f32 or None)."""
    lib = _capi.load()
    dev = a_planes_t.device
    total_rows = samples * out_rows if per_sample_tiles else out_rows
    d = _capi.ConvDesc()
    d.a = a_planes_t.data_ptr(); d.a_planes = a_planes_t.shape[0]
    d.samples = samples; d.a_rows = a_rows; d.a_ld = a_ld
    d.w = w_planes_t.data_ptr(); d.taps = taps; d.k_per_tap = k_per_tap; d.n_pad = n_pad
    d.per_sample_tiles = int(per_sample_tiles); d.tap_row_step = tap_row_step
    d.tap_col_step = tap_col_step; d.out_rows = out_rows; d.precision = precision
    keep = []
    if scale is not None:
        d.scale = scale.data_ptr(); d.shift = shift.data_ptr()
    d.relu = int(relu)
    if res is not None:
        d.res = res.data_ptr(); d.res_planes = res.shape[0]
        d.res_plane_stride = res[0].numel(); d.res_ld = res.shape[-1]
        d.res_rows_per_sample = res_rows_per_sample; d.res_row_step = res_row_step
        d.res_row_off = res_row_off; d.res_sample_div = res_sample_div
    out = out32 = None
    if out_f32_cols is None:
        out = torch.full((out_planes, total_rows, n_pad), float("nan"),
                         dtype=torch.float16 if precision == 3 else torch.bfloat16, device=dev)
        d.out = out.data_ptr(); d.out_planes = out_planes; d.out_plane_stride = out[0].numel()
        d.out_ld = n_pad
    else:
        out32 = torch.full((total_rows, out_f32_cols), float("nan"), dtype=torch.float32, device=dev)
        d.out_f32 = out32.data_ptr(); d.out_f32_ld = out_f32_cols; d.n_valid = out_f32_cols
    if stats is not None:
        d.stats = stats.data_ptr()
    stream = torch.cuda.current_stream().cuda_stream
    _capi.check(lib.vp3d_conv_gemm(ctypes.byref(d), stream), "vp3d_conv_gemm")
    torch.cuda.synchronize()
    return out, out32


def expected_conv(a_val, w_val, *, samples, a_rows, taps, k_per_tap, per_sample_tiles, tap_row_step,
                  tap_col_step, out_rows):
    """fp64 expectation of the raw accumulator.  a_val: [samples*a_rows, a_ld] fp64,
    w_val: [taps, n_pad, k_per_tap] fp64.  Returns [total_rows, n_pad]."""
    a_ld = a_val.shape[-1]
    a3 = a_val.reshape(samples, a_rows, a_ld)
    n_pad = w_val.shape[1]
    if per_sample_tiles:
        acc = torch.zeros(samples, out_rows, n_pad, dtype=torch.float64, device=a_val.device)
        for tap in range(taps):
            r0 = tap * tap_row_step
            c0 = tap * tap_col_step
            rows = a3[:, r0:r0 + out_rows, c0:c0 + k_per_tap]
            if rows.shape[1] < out_rows:  # TMA zero fill past the end of the sample
                pad = torch.zeros(samples, out_rows - rows.shape[1], k_per_tap, dtype=torch.float64,
                                  device=a_val.device)
                rows = torch.cat([rows, pad], dim=1)
            acc += rows @ w_val[tap].T
        return acc.reshape(samples * out_rows, n_pad)
    acc = torch.zeros(out_rows, n_pad, dtype=torch.float64, device=a_val.device)
    flat = a3.reshape(samples * a_rows, a_ld)
    for tap in range(taps):
        r0 = tap * tap_row_step
        c0 = tap * tap_col_step
        acc += flat[r0:r0 + out_rows, c0:c0 + k_per_tap] @ w_val[tap].T
    return acc
