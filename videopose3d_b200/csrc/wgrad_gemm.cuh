// Weight-gradient GEMM of the temporal convolutions (autograd backward of model.py:102,113-118,
// 167,178-180,33 — "conv bwd-filter"):
//
//   dW[tap][co][ci] = sum_{row} dZ[row, co] * X[rowmap(row, tap), ci]
//
// Both operands are channel-last activations, i.e. the reduction index (row) is the slow dimension
// of both: tcgen05 reads them as MN-major operands straight from the TMA-landed tiles, no
// transposed copies.  The reduction is split over `splits` row ranges; each split writes its own
// fp32 partial [split][tap][m_pad][n_pad] with plain coalesced stores (deterministic, no atomics);
// wgrad_reduce sums the partials into the (c_out, c_in, taps) fp32 layout of Conv1d.weight.grad.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vp3d {

struct WgradArgs {
  int per_sample;       // 0: rows flattened; 1: K loop runs per sample (dilated layout)
  int samples;
  int rows;             // per_sample: valid dZ rows per sample (L_out); flat: total rows
  int kchunks;          // 64-row chunks (per sample when per_sample)
  int taps;
  int tap_row_step;     // X row offset per tap (dilation), 0 for column-block taps
  int tap_col_step;     // X column offset per tap (strided layout), else 0
  int m_tiles;          // padded C_out / 128
  int n_tiles;          // padded C_in  / BLOCK_N
  int m_pad, n_pad;
  int pairs;            // 1 or 3 (bf16x3)
  int splits;
  float* partial;       // [splits][taps][m_pad][n_pad]
};

// Whether launch_wgrad_gemm runs this launch on CTA pairs (256 x 256 tiles over two SMs); the split
// count is best chosen for num_sms / 2 workers then.
bool wgrad_gemm_uses_pair(const WgradArgs& args, int block_n, int num_sms);

cudaError_t launch_wgrad_gemm(const CUtensorMap& tmap_dz, const CUtensorMap& tmap_x,
                              const WgradArgs& args, int block_n, int num_sms, cudaStream_t stream);

// grad[(co*c_in + ci)*taps_out + tap] = sum_s partial[s][tap_p][co][n]  with
//   merged == 0: tap_p = tap, n = ci                       (one slab per tap)
//   merged == 1: tap_p = 0,   n = tap*c_in + ci            (expand conv on the strided layout)
cudaError_t launch_wgrad_reduce(const float* partial, float* grad, int splits, int taps_p, int m_pad,
                                int n_pad, int c_out, int c_in, int taps_out, int merged,
                                cudaStream_t stream);

}  // namespace vp3d
