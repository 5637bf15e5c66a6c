// Bandwidth-bound helper kernels around the conv GEMMs: fp32 -> bf16 (hi/lo plane) packing of the
// network input and of the Conv1d weights, and folding of BatchNorm1d eval statistics into a
// per-channel affine.  Declarations only; definitions in pack.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace vp3d {

// x: fp32 (N, T, c_raw) contiguous (model.py:68 view of (N,T,J,F)).
// out: bf16 [planes][N][rows][k_pad]; row r of sample n gathers `group` consecutive frames starting
// at frame r*frame_step, i.e. columns [0, group*c_raw) = x[n, r*frame_step : r*frame_step+group, :]
// flattened, zero padded to k_pad.  (group = 1, frame_step = 1 for the dilated layout; group =
// frame_step = w0 for the strided layout where expand_conv becomes a plain GEMM.)
//
// Row order.  Default (perm == nullptr or perm->levels == 0): row index = n*rows + r.  Tap-major
// order (eval strided schedule): the rows of every activation are ordered so that the `w` taps of
// the next strided conv are `w` contiguous row regions.  With block widths w_1..w_B and
// R_i = N * rows_out(block i):  pos_0(n, t0) = (t0 mod w_1)*R_1 + pos_1(n, t0 / w_1), ...,
// pos_B(n, t) = n*rows_out(B) + t.  `perm` carries (R_i, w_i) for i = 1..levels and rows_out(B).
// f16 != 0 (both pack launchers): planes must be 1 and the 16-bit slots receive IEEE fp16 bits
// (saturated at +-65504) instead of bf16 -- the operand format of the fp16 eval mode.
struct PackPerm {
  int levels;        // number of residual blocks B (0 = natural order)
  int last_rows;     // rows per sample after the last block
  unsigned region[8];   // R_1 .. R_B  (row counts < 2^31)
  int width[8];         // w_1 .. w_B
};
cudaError_t launch_pack_input(const float* x, __nv_bfloat16* out, int planes, int N, int T,
                              int c_raw, int rows, int group, int frame_step, int k_pad,
                              long long plane_stride, cudaStream_t stream,
                              const PackPerm* perm = nullptr, int f16 = 0);

// w: fp32 Conv1d weight (c_out, c_in, taps) (tap index innermost, model.py:102,113-118).
// out: bf16 [planes][taps_out][n_pad][k_pad], zero padded.
//   merge_taps = 0: taps_out = taps, out[pl][tap][co][ci] = w[co][ci][tap]
//   merge_taps = 1: taps_out = 1,    out[pl][0][co][tap*c_in + ci] = w[co][ci][tap]
cudaError_t launch_pack_conv_weight(const float* w, __nv_bfloat16* out, int planes, int c_out,
                                    int c_in, int taps, int n_pad, int k_pad, int merge_taps,
                                    cudaStream_t stream, int f16 = 0);

// Eval-mode BatchNorm1d (model.py:32,117,119; eps = 1e-5) as y = x*scale + shift.
// gamma/beta/mean/var: [c]; scale/shift: [c_pad] (padding: scale 0, shift 0).
cudaError_t launch_bn_fold(const float* gamma, const float* beta, const float* mean,
                           const float* var, float eps, float* scale, float* shift, int c,
                           int c_pad, cudaStream_t stream);

// shrink bias (model.py:33): scale = 1, shift = bias, padded with zeros.
cudaError_t launch_bias_affine(const float* bias, float* scale, float* shift, int c, int c_pad,
                               cudaStream_t stream);

}  // namespace vp3d
