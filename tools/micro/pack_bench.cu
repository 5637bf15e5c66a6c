// Micro-benchmark of the input pack (configs[1] shape) against plain streaming kernels, to tell what
// bounds it.  Build: make micro   Run on the GPU box: videopose3d_b200/_lib/dbg/pack_bench
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../videopose3d_b200/csrc/pack.cuh"

using namespace vp3d;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void stream_copy(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(in + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    out[i] = o;
  }
}
__global__ void read_only(const float4* __restrict__ in, float* __restrict__ out, long long n4) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(in + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) out[0] = acc;
}
__global__ void write_only(uint4* __restrict__ out, long long n16) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = make_uint4(1, 2, 3, 4);
}

int main() {
  const int N = 1024, T = 243, C = 34, rows = 81, group = 3, k_pad = 128;
  const long long in_f = (long long)N * T * C, out_e = (long long)N * rows * k_pad;
  float* x; __nv_bfloat16* out; uint8_t* flush;
  CK(cudaMalloc(&x, in_f * 4)); CK(cudaMalloc(&out, out_e * 2)); CK(cudaMalloc(&flush, 256 << 20));
  CK(cudaMemset(x, 0, in_f * 4));
  PackPerm perm; memset(&perm, 0, sizeof(perm));
  perm.levels = 4; perm.last_rows = 1;
  const unsigned reg[4] = {27648u, 9216u, 3072u, 1024u};
  for (int i = 0; i < 4; ++i) { perm.region[i] = reg[i]; perm.width[i] = 3; }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const char* names[] = {"pack tap-major f16", "pack natural f16", "stream copy fp32->bf16 (34MB->17MB)",
                         "read only 34MB", "write only 21MB", "pack tap-major, warm L2"};
  for (int v = 0; v < 6; ++v) {
    float best = 1e9f, sum = 0.f;
    const int iters = 12;
    for (int it = 0; it < iters; ++it) {
      if (v != 5) CK(cudaMemsetAsync(flush, it, 256 << 20));
      CK(cudaEventRecord(e0));
      switch (v) {
        case 0: case 5: CK(launch_pack_input(x, out, 1, N, T, C, rows, group, group, k_pad, 0, 0, &perm, 1)); break;
        case 1: CK(launch_pack_input(x, out, 1, N, T, C, rows, group, group, k_pad, 0, 0, nullptr, 1)); break;
        case 2: stream_copy<<<148 * 16, 256>>>((const float4*)x, (uint2*)out, in_f / 4); break;
        case 3: read_only<<<148 * 16, 256>>>((const float4*)x, (float*)out, in_f / 4); break;
        case 4: write_only<<<148 * 16, 256>>>((uint4*)out, out_e * 2 / 16); break;
      }
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    printf("%-40s best %.2f us  mean %.2f us\n", names[v], best * 1e3f, sum / (iters - 2) * 1e3f);
  }
  return 0;
}
