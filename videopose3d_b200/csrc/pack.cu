#include "pack.cuh"

#include <cuda_fp16.h>

#include <string.h>

#include <stdint.h>

namespace vp3d {

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// 16-bit storage of the fp16 eval mode: IEEE half bits in a bf16-typed slot (single plane),
// saturating at +-65504 like the GEMM epilogue does.
__device__ __forceinline__ __nv_bfloat16 f16_bits(float v) {
  v = fminf(fmaxf(v, -65504.0f), 65504.0f);
  const unsigned short h = __half_as_ushort(__float2half_rn(v));
  return __ushort_as_bfloat16(h);
}

// One thread per (row, 8-column group): 16-byte stores, coalesced along the row.
__global__ void pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                  int planes, int N, int T, int c_raw, int rows, int group,
                                  int frame_step, int k_pad, long long plane_stride,
                                  const PackPerm perm, int f16) {
  const int groups_per_row = k_pad >> 3;
  const long long total = (long long)N * rows * groups_per_row;
  const int k_valid = group * c_raw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int g;
    long long row;
    if (total < 0x7fffffffll) {  // 32-bit index math (the common case): far cheaper divisions
      const unsigned iu = (unsigned)i;
      const unsigned ru = iu / (unsigned)groups_per_row;
      g = (int)(iu - ru * (unsigned)groups_per_row);
      row = ru;
    } else {
      g = (int)(i % groups_per_row);
      row = i / groups_per_row;
    }
    int r, n;
    if (perm.levels == 0) {
      if (total < 0x7fffffffll) {
        n = (int)((unsigned)row / (unsigned)rows);
        r = (int)((unsigned)row - (unsigned)n * (unsigned)rows);
      } else {
        r = (int)(row % rows);
        n = (int)(row / rows);
      }
    } else {
      // tap-major order (row < 2^31): peel one tap digit per block, outermost first; digit i has
      // weight w_1 * ... * w_(i-1) in the frame index, the in-sample row of the last block the
      // product of all widths
      unsigned j = (unsigned)row;
      int frame = 0, weight = 1;
      for (int lv = 0; lv < perm.levels; ++lv) {
        const unsigned dgt = j / perm.region[lv];
        j -= dgt * perm.region[lv];
        frame += (int)dgt * weight;
        weight *= perm.width[lv];
      }
      n = (int)(j / (unsigned)perm.last_rows);
      r = frame + (int)(j - (unsigned)n * (unsigned)perm.last_rows) * weight;
    }
    const float* src = x + ((long long)n * T + (long long)r * frame_step) * c_raw;
    __align__(16) __nv_bfloat16 hi[8];
    __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      const float v = (k < k_valid) ? __ldg(src + k) : 0.0f;
      if (f16) hi[j] = f16_bits(v); else split_bf16(v, hi[j], lo[j]);
    }
    __nv_bfloat16* dst = out + row * k_pad + g * 8;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(hi);
    if (planes == 2) *reinterpret_cast<uint4*>(dst + plane_stride) = *reinterpret_cast<const uint4*>(lo);
  }
}

cudaError_t launch_pack_input(const float* x, __nv_bfloat16* out, int planes, int N, int T,
                              int c_raw, int rows, int group, int frame_step, int k_pad,
                              long long plane_stride, cudaStream_t stream, const PackPerm* perm,
                              int f16) {
  PackPerm pp;
  memset(&pp, 0, sizeof(pp));
  if (perm) pp = *perm;
  const long long total = (long long)N * rows * (k_pad >> 3);
  if (total <= 0) return cudaSuccess;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_input_kernel<<<(int)blocks, threads, 0, stream>>>(x, out, planes, N, T, c_raw, rows, group,
                                                         frame_step, k_pad, plane_stride, pp, f16);
  return cudaGetLastError();
}

// One thread per (co, k) position of the padded slab; the thread walks the taps, so the fp32 reads of
// a warp cover one contiguous span of Conv1d.weight and each tap slab receives a coalesced bf16 row.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                        int planes, int c_out, int c_in, int taps, int n_pad,
                                        int k_pad, int merged, int f16) {
  const long long slab = (long long)n_pad * k_pad;
  const long long plane_elems = (merged ? 1 : taps) * slab;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < slab;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % k_pad);
    const int co = (int)(i / k_pad);
    if (merged) {
      float v = 0.0f;
      if (co < c_out && k < taps * c_in) {
        const int tap = k / c_in, ci = k - tap * c_in;
        v = __ldg(w + ((long long)co * c_in + ci) * taps + tap);
      }
      __nv_bfloat16 hi, lo;
      if (f16) { hi = f16_bits(v); lo = hi; } else split_bf16(v, hi, lo);
      out[i] = hi;
      if (planes == 2) out[plane_elems + i] = lo;
    } else {
      const bool in = co < c_out && k < c_in;
      const float* src = w + ((long long)co * c_in + k) * taps;
      for (int tap = 0; tap < taps; ++tap) {
        const float v = in ? __ldg(src + tap) : 0.0f;
        __nv_bfloat16 hi, lo;
        if (f16) { hi = f16_bits(v); lo = hi; } else split_bf16(v, hi, lo);
        out[tap * slab + i] = hi;
        if (planes == 2) out[plane_elems + tap * slab + i] = lo;
      }
    }
  }
}

cudaError_t launch_pack_conv_weight(const float* w, __nv_bfloat16* out, int planes, int c_out,
                                    int c_in, int taps, int n_pad, int k_pad, int merge_taps,
                                    cudaStream_t stream, int f16) {
  const long long total = (long long)n_pad * k_pad;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_conv_weight_kernel<<<(int)blocks, threads, 0, stream>>>(w, out, planes, c_out, c_in, taps,
                                                               n_pad, k_pad, merge_taps, f16);
  return cudaGetLastError();
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, float* __restrict__ scale, float* __restrict__ shift,
                               int c, int c_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c_pad) return;
  float s = 0.0f, b = 0.0f;
  if (i < c) {
    s = gamma[i] / sqrtf(var[i] + eps);
    b = beta[i] - mean[i] * s;
  }
  scale[i] = s;
  shift[i] = b;
}

cudaError_t launch_bn_fold(const float* gamma, const float* beta, const float* mean,
                           const float* var, float eps, float* scale, float* shift, int c,
                           int c_pad, cudaStream_t stream) {
  bn_fold_kernel<<<(c_pad + 255) / 256, 256, 0, stream>>>(gamma, beta, mean, var, eps, scale, shift,
                                                          c, c_pad);
  return cudaGetLastError();
}

__global__ void bias_affine_kernel(const float* __restrict__ bias, float* __restrict__ scale,
                                   float* __restrict__ shift, int c, int c_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c_pad) return;
  scale[i] = (i < c) ? 1.0f : 0.0f;
  shift[i] = (i < c) ? bias[i] : 0.0f;
}

cudaError_t launch_bias_affine(const float* bias, float* scale, float* shift, int c, int c_pad,
                               cudaStream_t stream) {
  bias_affine_kernel<<<(c_pad + 255) / 256, 256, 0, stream>>>(bias, scale, shift, c, c_pad);
  return cudaGetLastError();
}

}  // namespace vp3d
