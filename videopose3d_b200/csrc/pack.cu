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

// Rows are walked in SOURCE order (sample, then frame group) so that the fp32 read is one
// sequential stream; the scattered side (tap-major row order) is the 16-bit output, which stays in
// L2 for the expand GEMM.  A block owns a chunk of at most kPackRows consecutive rows of one sample:
// the row -> output-row map of the chunk (up to eight integer divisions per row) is computed once
// into shared memory, then every warp converts whole rows -- lanes stride over the element pairs
// with 8-byte loads and 4-byte stores, two rows in flight per warp.
constexpr int kPackRows = 96;

template <bool VEC2>
__device__ __forceinline__ void pack_load_pair(const float* src, int k, int k_valid, float& v0,
                                               float& v1) {
  v0 = 0.0f;
  v1 = 0.0f;
  if (VEC2) {   // c_raw even and x 8-byte aligned: k_valid is even too
    if (k < k_valid) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(src + k));
      v0 = v.x;
      v1 = v.y;
    }
  } else {
    if (k < k_valid) v0 = __ldg(src + k);
    if (k + 1 < k_valid) v1 = __ldg(src + k + 1);
  }
}

__device__ __forceinline__ void pack_store_pair(__nv_bfloat16* dst, long long plane_stride, int planes,
                                                int f16, float v0, float v1) {
  __nv_bfloat162 hi, lo;
  if (f16) {
    hi.x = f16_bits(v0);
    hi.y = f16_bits(v1);
  } else {
    split_bf16(v0, hi.x, lo.x);
    split_bf16(v1, hi.y, lo.y);
  }
  *reinterpret_cast<__nv_bfloat162*>(dst) = hi;
  if (planes == 2) *reinterpret_cast<__nv_bfloat162*>(dst + plane_stride) = lo;
}

template <bool VEC2>
__global__ void __launch_bounds__(256)
pack_input_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int planes, int N,
                  int T, int c_raw, int rows, int group, int frame_step, int k_pad,
                  long long plane_stride, const PackPerm perm, int f16) {
  __shared__ long long s_row[kPackRows];
  const int chunks = (rows + kPackRows - 1) / kPackRows;
  const long long items = (long long)N * chunks;
  const int k_valid = group * c_raw;
  const int pairs = k_pad >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const long long n = item / chunks;
    const int r0 = (int)(item - n * chunks) * kPackRows;
    const int nr = min(kPackRows, rows - r0);
    __syncthreads();   // the previous chunk's map is no longer read
    if ((int)threadIdx.x < nr) {
      const int r = r0 + (int)threadIdx.x;
      // output row: natural order n*rows + r, or the tap-major position (pack.cuh): peel one tap
      // digit per block, innermost frame digit first
      long long row;
      if (perm.levels == 0) {
        row = n * rows + r;
      } else {
        unsigned t = (unsigned)r;
        row = 0;
        for (int lv = 0; lv < perm.levels; ++lv) {
          const unsigned w = (unsigned)perm.width[lv];
          const unsigned q = t / w;
          row += (long long)(t - q * w) * perm.region[lv];
          t = q;
        }
        row += n * perm.last_rows + t;
      }
      s_row[threadIdx.x] = row;
    }
    __syncthreads();
    const float* src0 = x + (n * T + (long long)r0 * frame_step) * c_raw;
    const long long src_step = (long long)frame_step * c_raw;
    for (int rl = warp * 2; rl < nr; rl += 16) {   // 8 warps x 2 rows
      const bool two = rl + 1 < nr;
      const float* sa = src0 + rl * src_step;
      const float* sb = sa + src_step;
      __nv_bfloat16* da = out + s_row[rl] * k_pad;
      __nv_bfloat16* db = out + s_row[two ? rl + 1 : rl] * k_pad;
      for (int c = lane; c < pairs; c += 64) {
        const int c2 = c + 32;
        float a0, a1, a2, a3, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
        pack_load_pair<VEC2>(sa, 2 * c, k_valid, a0, a1);
        pack_load_pair<VEC2>(sa, 2 * c2, c2 < pairs ? k_valid : 0, a2, a3);
        if (two) {
          pack_load_pair<VEC2>(sb, 2 * c, k_valid, b0, b1);
          pack_load_pair<VEC2>(sb, 2 * c2, c2 < pairs ? k_valid : 0, b2, b3);
        }
        pack_store_pair(da + 2 * c, plane_stride, planes, f16, a0, a1);
        if (c2 < pairs) pack_store_pair(da + 2 * c2, plane_stride, planes, f16, a2, a3);
        if (two) {
          pack_store_pair(db + 2 * c, plane_stride, planes, f16, b0, b1);
          if (c2 < pairs) pack_store_pair(db + 2 * c2, plane_stride, planes, f16, b2, b3);
        }
      }
    }
  }
}

cudaError_t launch_pack_input(const float* x, __nv_bfloat16* out, int planes, int N, int T,
                              int c_raw, int rows, int group, int frame_step, int k_pad,
                              long long plane_stride, cudaStream_t stream, const PackPerm* perm,
                              int f16) {
  PackPerm pp;
  memset(&pp, 0, sizeof(pp));
  if (perm) pp = *perm;
  if (k_pad % 8) return cudaErrorInvalidValue;
  if (N <= 0 || rows <= 0) return cudaSuccess;
  const long long items = (long long)N * ((rows + kPackRows - 1) / kPackRows);
  long long blocks = items;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const bool vec2 = (c_raw % 2 == 0) && (reinterpret_cast<uintptr_t>(x) % 8 == 0);
  // (a plain launch: the input pack is the first kernel of a forward and usually follows a copy or
  // an event wait, where a programmatic edge buys nothing)
  if (vec2)
    pack_input_kernel<true><<<(int)blocks, 256, 0, stream>>>(
        x, out, planes, N, T, c_raw, rows, group, frame_step, k_pad, plane_stride, pp, f16);
  else
    pack_input_kernel<false><<<(int)blocks, 256, 0, stream>>>(
        x, out, planes, N, T, c_raw, rows, group, frame_step, k_pad, plane_stride, pp, f16);
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
